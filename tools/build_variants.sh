#!/bin/bash
# Builds tuning variants of libcsa_b200.so (same ABI, different -D knobs of attn_sm100.cu) into spider_b200/variants/.
# usage: tools/build_variants.sh "name1 -DCSA_TOKEN_CHUNK=1" "name2 -DCSA_POLY_PAIRS=2 -DCSA_PINGPONG=0" "trace -DCSA_TRACE=1" ...
# then on the GPU box:  bash tools/sweep.sh tag   (times every variant with tools/bench_kernel.py)
set -e
cd "$(dirname "$0")/../spider_b200/csrc"
make -j4 > /dev/null          # the other translation units are shared with the in-tree build
mkdir -p ../variants build
ARCH="-gencode arch=compute_100a,code=sm_100a"
for v in "$@"; do
  set -- $v
  name=$1
  shift
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -Xptxas -v \
    -c attn_sm100.cu -o build/attn_$name.o 2>&1 | grep -i "spill\|error" || true
  nvcc $ARCH -shared -o ../variants/libcsa_$name.so build/abi.o build/compact.o build/peer.o build/linear.o build/gemm_sm100.o \
    build/attn_$name.o -lcudart -lcublasLt -Xlinker -rpath=/usr/local/cuda/lib64
  echo built $name
done
