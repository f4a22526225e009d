mkdir -p gpurun_out/r03t
for f in "" "--projections cublaslt" "" "--projections cublaslt"; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-hbm --no-config4 $f > gpurun_out/r03t/b.json 2>gpurun_out/r03t/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/r03t/b.json').read().strip().splitlines()[-1]);print('[$f]',round(d['value'],2),round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['roofline']['achieved'])"
done
