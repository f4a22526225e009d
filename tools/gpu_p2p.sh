#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_p2p.sh tag N "16 4"' : multi-GPU parity check, then bench.py with both exchange paths
TAG=${1:-p2p}; N=${2:-4}; FRAMES=${3:-16}; EXS=${4:-"p2p nccl"}
mkdir -p gpurun_out/$TAG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/gpu_dist_check.py > gpurun_out/$TAG/check_n$N.log 2>&1
echo "dist check exit $?"; tail -4 gpurun_out/$TAG/check_n$N.log
for F in $FRAMES; do
 for EX in $EXS; do
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --frames $F --no-cpu --exchange $EX > gpurun_out/$TAG/bench_f${F}_n${N}_$EX.json 2> gpurun_out/$TAG/bench_f${F}_n${N}_$EX.err
  echo "F=$F N=$N $EX exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$TAG/bench_f${F}_n${N}_$EX.json").read().strip().splitlines()[-1])
    print("  value",d["value"],"ms/step",d["ms_per_step"],"kernel",d["roofline"]["achieved"],"attn share",d["roofline"]["attn_share_of_step"],"host ms",d.get("host_issue_ms_per_step"),"e2e",d.get("e2e",{}).get("value"))
except Exception as e:
    print("  no json:",e); print(open("gpurun_out/$TAG/bench_f${F}_n${N}_$EX.err").read()[-2500:])
PY
 done
done
