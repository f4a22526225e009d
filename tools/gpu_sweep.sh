#!/bin/bash
# BASELINE config 5 (roofline / scaling sweep): bench.py over sa x frames x resolution at N GPUs.
# usage: gpurun [--gpus N] --timeout 1500 -- 'bash tools/gpu_sweep.sh tag N "0.0 0.25 0.5 1.0" "2 4 8 16" "768 1024 1536"'
# One JSON line per point under gpurun_out/<tag>/; a table of (sa, F, res) -> ms/step, TFLOP/s, kernel TFLOP/s on stdout.
TAG=${1:-sweep}; N=${2:-1}; SAS=${3:-"0.0 0.5 1.0"}; FRAMES=${4:-"2 4 8"}; RES=${5:-"768 1024"}
mkdir -p gpurun_out/$TAG
if [ "$N" = "1" ]; then RUN="python"; else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"; fi
echo "sa frames res n_gpus ms_per_step tflops kernel_tflops"
for R in $RES; do for F in $FRAMES; do for SA in $SAS; do
  # frames must split over the N/2 ranks of a CFG half
  if [ "$N" -gt 2 ] && [ $((F % (N / 2))) -ne 0 ]; then continue; fi
  OUT=gpurun_out/$TAG/bench_sa${SA}_f${F}_r${R}_n${N}.json
  timeout 300 $RUN bench.py --gpus $N --steps 5 --warmup 3 --frames $F --res $R --sa $SA --no-cpu --no-e2e > $OUT 2> ${OUT%.json}.err
  python - "$OUT" "$SA" "$F" "$R" "$N" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], d["ms_per_step"], d["value"], d["roofline"]["achieved"])
except Exception as e:   # noqa: BLE001
    print(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], "failed:", e)
PY
done; done; done
