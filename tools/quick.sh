#!/bin/bash
# gpurun -- 'bash tools/quick.sh tag' : kernel-only timing of the in-tree build + GPU parity tests
TAG=${1:-q}; mkdir -p gpurun_out/$TAG
timeout 300 python tools/bench_kernel.py intree 2>&1 | tee gpurun_out/$TAG/kernel.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/$TAG/pytest.log
