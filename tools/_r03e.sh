mkdir -p gpurun_out/r03e
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_native_layer.py tests/test_gpu_parity_baseline.py tests/test_gpu_graph.py -m gpu -x -q > gpurun_out/r03e/pytest.log 2>&1; tail -3 gpurun_out/r03e/pytest.log
for f in "" "--projections cublaslt"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-hbm --no-config4 $f > gpurun_out/r03e/b.json 2>gpurun_out/r03e/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/r03e/b.json').read().strip().splitlines()[-1]);print('[$f]',round(d['value'],2),round(d['ms_per_step'],4))"
done
