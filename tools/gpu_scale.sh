#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_scale.sh tag "1 2 4" [frames...]' : bench.py at several rank counts
TAG=${1:-scale}; NS=${2:-"1 2"}; shift; shift
FRAMES=${*:-4}
mkdir -p gpurun_out/$TAG
for F in $FRAMES; do
 for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --frames $F --no-cpu > gpurun_out/$TAG/bench_f${F}_n$N.json 2> gpurun_out/$TAG/bench_f${F}_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --frames $F --no-cpu > gpurun_out/$TAG/bench_f${F}_n$N.json 2> gpurun_out/$TAG/bench_f${F}_n$N.err
  fi
  echo "F=$F N=$N exit $?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$TAG/bench_f${F}_n$N.json").read().strip().splitlines()[-1])
    print("  value",d["value"],"ms/step",d["ms_per_step"],"kernel",d["roofline"]["achieved"],"e2e",d.get("e2e",{}).get("value"))
except Exception as e:
    print("  no json:",e); print(open("gpurun_out/$TAG/bench_f${F}_n$N.err").read()[-1500:])
PY
 done
done
