#!/bin/bash
# gpurun -- 'bash tools/prof_variant.sh <variant .so> <tag>' : ncu --set full of the 64x64-layer attention launch
LIB=$1; TAG=${2:-pv}
mkdir -p gpurun_out/$TAG
CSA_B200_LIB=$PWD/$LIB timeout 600 ncu --set full --clock-control none --import-source on -k regex:csa_attn_kernel -s 3 -c 1 \
  -o gpurun_out/$TAG/attn -f python tools/bench_kernel.py > gpurun_out/$TAG/log.txt 2>&1
tail -3 gpurun_out/$TAG/log.txt
