#!/bin/bash
# gpurun -- 'bash tools/sweep.sh tag'  : time every variant under spider_b200/variants/ (trace builds: timeline instead)
TAG=${1:-sweep}
mkdir -p gpurun_out/$TAG
for lib in spider_b200/variants/libcsa_*.so; do
  name=$(basename $lib .so)
  case $name in
    *trace*) CSA_B200_LIB=$PWD/$lib timeout 120 python tools/trace_timeline.py 4096 640 10 > gpurun_out/$TAG/timeline_$name.log 2>&1
             head -30 gpurun_out/$TAG/timeline_$name.log ;;
    *)       CSA_B200_LIB=$PWD/$lib timeout 120 python tools/bench_kernel.py ${name#libcsa_} 2>&1 | tee -a gpurun_out/$TAG/sweep.log ;;
  esac
done
