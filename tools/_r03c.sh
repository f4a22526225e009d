mkdir -p gpurun_out/r03c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03c/pytest.log 2>&1; tail -3 gpurun_out/r03c/pytest.log
for v in old new; do
  if [ $v = old ]; then export CSA_B200_LIB=spider_b200/variants/libcsa_old.so; else unset CSA_B200_LIB; fi
  timeout 300 python tools/bench_kernel.py $v 2>&1 | tail -3
done
unset CSA_B200_LIB
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-hbm --no-config4 > gpurun_out/r03c/b.json 2>gpurun_out/r03c/b.err
python -c "
import json;d=json.loads(open('gpurun_out/r03c/b.json').read().strip().splitlines()[-1]);print(round(d['value'],2),round(d['ms_per_step'],4), d['roofline']['achieved'], d['roofline']['frac'])"
