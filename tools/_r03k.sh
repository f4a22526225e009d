mkdir -p gpurun_out/r03k
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r03k/pytest.log 2>&1; tail -3 gpurun_out/r03k/pytest.log
for v in prev new prev new; do
  if [ $v = prev ]; then export CSA_B200_LIB=spider_b200/variants/libcsa_prev.so; else unset CSA_B200_LIB; fi
  timeout 300 python tools/bench_kernel.py $v 2>&1 | tail -3
done
