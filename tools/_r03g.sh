mkdir -p gpurun_out/r03g
W=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29533 tools/gpu_dist_check.py --steps 20 --skew --graph > gpurun_out/r03g/check_w$W.log 2>&1; echo "check exit $?"; tail -4 gpurun_out/r03g/check_w$W.log
for fe in 1 0; do
CSA_FUSED_EXCHANGE=$fe timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $W --steps 10 --warmup 3 --no-cpu --no-e2e --no-hbm > gpurun_out/r03g/bench_w${W}_fe$fe.json 2> gpurun_out/r03g/bench_w${W}_fe$fe.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r03g/bench_w${W}_fe$fe.json').read().strip().splitlines()[-1])
    print('fused=$fe N=$W', d['value'], d['ms_per_step'], 'config4', d.get('config4',{}).get('ms_per_step'), d.get('config4',{}).get('efficiency_vs_n1_ms'), d['launches_by_entry'])
except Exception as e:
    print('no json', e); print(open('gpurun_out/r03g/bench_w${W}_fe$fe.err').read()[-1500:])
PY
done
