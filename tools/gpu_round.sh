#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list and one full capture of the attention
# kernel.  Everything lands in gpurun_out/ (scratch); summaries worth keeping are copied to profiles/ by hand.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag] [what...]'   what = tests bench launches full check
set -u
TAG=${1:-r01}
shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > "$OUT/gpu.csv" 2>&1

for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
      echo "pytest exit $?" | tee -a "$OUT/pytest_gpu.log"
      tail -5 "$OUT/pytest_gpu.log"
      ;;
    smoke)
      timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > "$OUT/smoke.log" 2>&1
      echo "smoke exit $?"; tail -4 "$OUT/smoke.log"
      ;;
    check)
      timeout 600 python tools/gpu_check.py > "$OUT/gpu_check.log" 2>&1
      echo "gpu_check exit $?"; tail -40 "$OUT/gpu_check.log"
      ;;
    bench)
      timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
      echo "bench exit $?"; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"
      ;;
    benchref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
      echo "bench ref exit $?"; cat "$OUT/bench_ref.json"
      ;;
    launches)
      # every launch of one warm step with its device time (cold-cache, serialised: compare SHARES)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
        --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-graph --no-hbm > "$OUT/launches.log" 2>&1
      echo "ncu launches exit $?"
      ;;
    f16)
      timeout 600 python bench.py --frames 16 --steps 5 --warmup 3 --no-cpu --no-e2e --no-hbm --share-weights > "$OUT/bench_f16_n1.json" 2> "$OUT/bench_f16_n1.err"
      echo "bench F=16 exit $?"; head -c 600 "$OUT/bench_f16_n1.json"; tail -3 "$OUT/bench_f16_n1.err"
      ;;
    all70)
      timeout 600 python bench.py --placement all --steps 5 --warmup 3 --no-cpu --no-hbm > "$OUT/bench_all70.json" 2> "$OUT/bench_all70.err"
      echo "bench all-70 exit $?"; head -c 600 "$OUT/bench_all70.json"; tail -3 "$OUT/bench_all70.err"
      ;;
    sweep)
      timeout 1200 python tools/sweep_inproc.py --out "$OUT/sweep_n1.md" > "$OUT/sweep_n1.jsonl" 2> "$OUT/sweep_n1.err"
      echo "sweep exit $?"; tail -60 "$OUT/sweep_n1.md"; tail -3 "$OUT/sweep_n1.err"
      ;;
    sdpa)
      # same-box GPU comparator: the reference call through torch SDPA with the dense mask
      timeout 600 python tools/bench_torch_sdpa.py > "$OUT/torch_sdpa.json" 2> "$OUT/torch_sdpa.err"
      echo "torch sdpa exit $?"; cat "$OUT/torch_sdpa.json"
      ;;
    full)
      # one 32x32-class and one 64x64-class attention launch of a warm step
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:csa_attn_kernel -s 136 -c 4 \
        -o "$OUT/attn_full" -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-graph --no-hbm > "$OUT/full.log" 2>&1
      echo "ncu full exit $?"; tail -3 "$OUT/full.log"
      ;;
  esac
done
ls -la "$OUT"
