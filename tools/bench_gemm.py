#!/usr/bin/env python
"""Hand-written projection GEMM (csa_gemm) vs cuBLASLt (csa_linear) on the shapes of the path, one B200.
Operands rotate over enough buffers to exceed the 126 MB L2; CUDA events on the launching stream.
    python tools/bench_gemm.py [--dtype bf16]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", choices=["bf16", "fp16"], default="bf16")
    args = ap.parse_args()
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    peak = 1649.1
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["bf16_tflops"]
    except (OSError, ValueError, KeyError):
        pass
    shapes = [("32x32 q / out  F=4", 8192, 1280, 1280, False), ("32x32 k|v      F=4", 8192, 2560, 1280, False),
              ("64x64 q / out  F=4", 32768, 640, 640, True), ("64x64 k|v      F=4", 32768, 1280, 640, False),
              ("32x32 k|v      F=16", 32768, 2560, 1280, False), ("64x64 k|v      F=16", 131072, 1280, 640, False),
              ("32x32 q  1 frame (N=8)", 1024, 1280, 1280, False), ("32x32 k|v read", 2048, 2560, 1280, False),
              ("32x32 q|k|v   F=4", 8192, 3840, 1280, False), ("64x64 q|k|v   F=4", 32768, 1920, 640, False),
              ("32x32 q|k|v 2 frames (N=4)", 2048, 3840, 1280, False),
              ("32x32 q|k|v 1 frame (N=8)", 1024, 3840, 1280, False)]
    rows = []
    for name, m, n, k, bias in shapes:
        rot = max(2, int(300e6 // ((m * k + m * n) * 2)) + 1)
        g = torch.Generator(device=dev).manual_seed(0)
        xs = [torch.randn((m, k), device=dev, generator=g).to(dtype) for _ in range(rot)]
        ys = [torch.empty((m, n), device=dev, dtype=dtype) for _ in range(rot)]
        w = (torch.randn((n, k), device=dev, generator=g) * k ** -0.5).to(dtype)
        b = torch.randn((n,), device=dev, generator=g).to(dtype) if bias else None
        t_own = timeit(lambda i: native.gemm(xs[i % rot], w, b, out=ys[i % rot]))
        t_lib = timeit(lambda i: native.linear(xs[i % rot], w, b, out=ys[i % rot]))
        fl = 2.0 * m * n * k
        row = {"shape": name, "m": m, "n": n, "k": k, "csa_gemm_us": round(t_own * 1e3, 2),
               "csa_gemm_tflops": round(fl / t_own * 1e-9, 1), "frac_of_peak": round(fl / t_own * 1e-9 / peak, 3),
               "cublaslt_us": round(t_lib * 1e3, 2), "cublaslt_tflops": round(fl / t_lib * 1e-9, 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del xs, ys
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
