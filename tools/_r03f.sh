mkdir -p gpurun_out/r03f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03f/pytest.log 2>&1; tail -3 gpurun_out/r03f/pytest.log
for e in 1 0 1 0; do
  CSA_PDL=$e timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-hbm --no-config4 > gpurun_out/r03f/b.json 2>gpurun_out/r03f/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/r03f/b.json').read().strip().splitlines()[-1]);print('[pdl=$e]',round(d['value'],2),round(d['ms_per_step'],4))"
done
