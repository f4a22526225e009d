#!/bin/bash
# Builds tuning variants of libcsa_b200.so (same ABI, different -D knobs) into spider_b200/variants/.
# usage: tools/build_variants.sh "PIPE POLY [SPLIT [PINGPONG [EXTRA -D flags]]]" ...   e.g.  tools/build_variants.sh "0 0" "0 4 2" "1 4"
set -e
cd "$(dirname "$0")/../spider_b200/csrc"
mkdir -p ../variants build
ARCH="-gencode arch=compute_100a,code=sm_100a"
for v in "$@"; do
  set -- $v
  pipe=$1
  poly=$2
  split=${3:-1}
  pp=${4:-1}
  shift 4 2>/dev/null || shift $#
  extra="$*"
  name="p${pipe}_k${poly}_s${split}_g$pp$(echo "$extra" | tr -d ' =-' | tr 'A-Z' 'a-z')"
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DCSA_SOFTMAX_PIPE=$pipe -DCSA_POLY_PAIRS=$poly -DCSA_ROW_SPLIT=$split -DCSA_PINGPONG=$pp $extra -Xptxas -v \
    -c attn_sm100.cu -o build/attn_$name.o
  nvcc $ARCH -shared -o ../variants/libcsa_$name.so build/abi.o build/compact.o build/attn_$name.o -lcudart
  echo built $name
done
