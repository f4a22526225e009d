#!/bin/bash
# gpurun -- 'bash tools/sweep.sh tag'  : time every variant under spider_b200/variants/
TAG=${1:-sweep}
mkdir -p gpurun_out/$TAG
for lib in spider_b200/variants/libcsa_*.so; do
  CSA_B200_LIB=$PWD/$lib timeout 120 python tools/bench_kernel.py 2>&1 | tee -a gpurun_out/$TAG/sweep.log
done
