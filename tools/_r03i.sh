W=${1:-8}
mkdir -p gpurun_out/r03i
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29533 tools/gpu_dist_check.py --steps 50 --skew --graph > gpurun_out/r03i/check_w$W.log 2>&1; echo "check exit $?"; tail -3 gpurun_out/r03i/check_w$W.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $W --steps 10 --warmup 3 > gpurun_out/r03i/bench_w$W.json 2> gpurun_out/r03i/bench_w$W.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r03i/bench_w$W.json').read().strip().splitlines()[-1])
    print('N=$W', d['value'], d['ms_per_step'], 'e2e', d.get('e2e',{}).get('ms_per_step'), 'config4', d.get('config4',{}).get('ms_per_step'), d.get('config4',{}).get('efficiency_vs_n1_ms'), d['roofline']['avg_launch_ms'], d['launches_by_entry'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/r03i/bench_w$W.err').read()[-1500:])
PY
