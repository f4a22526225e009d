mkdir -p gpurun_out/r03d
for d in 3 6; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-hbm --no-config4 --e2e-depth $d > gpurun_out/r03d/b$d.json 2>gpurun_out/r03d/b$d.err
python -c "
import json;d=json.loads(open('gpurun_out/r03d/b$d.json').read().strip().splitlines()[-1]);print($d, round(d['value'],2),round(d['ms_per_step'],4), d['e2e'])"
done
