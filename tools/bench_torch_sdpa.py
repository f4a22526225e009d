#!/usr/bin/env python
"""Same-box GPU comparator: what the REFERENCE's attention call costs on this B200 when torch runs it
— `F.scaled_dot_product_attention(q, k, v, attn_mask=dense_bool)` over all frames' tokens of a CFG half
(StoryDiffusion/Comic_Generation.py:175-177) — per SDPA backend, on the two SDXL layer classes of the bench workload.
Reported next to our kernel's numbers; library kernels, not part of the product path.

    python tools/bench_torch_sdpa.py [--frames 4] [--res 1024] [--sa 0.5] [--dtype bf16]
Prints one JSON line: per layer class and backend ms per call and ALGORITHMIC TFLOP/s (the same FLOP count bench.py
uses: only the keys the mask keeps), and the 30 + 6 layer step equivalent of the fastest backend."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--sa", type=float, default=0.5)
    ap.add_argument("--dtype", choices=["bf16", "fp16"], default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    Fl, T = args.frames, args.frames + 1
    from torch.nn.attention import SDPBackend, sdpa_kernel
    backends = {"efficient": SDPBackend.EFFICIENT_ATTENTION, "cudnn": SDPBackend.CUDNN_ATTENTION,
                "math": SDPBackend.MATH}
    out = {"what": "reference attention call through torch SDPA with the dense bool mask, same GPU",
           "frames": Fl, "res": args.res, "sa": args.sa, "dtype": args.dtype, "layers": {}}
    step_ms = {}
    classes = [("32x32", (args.res // 32) ** 2, 1280, 20, 30), ("64x64", (args.res // 16) ** 2, 640, 10, 6)]
    for name, N, C, heads, count in classes:
        torch.manual_seed(0)
        sample = torch.rand((T * N,), device=dev) < args.sa
        cm = csa_masks.CompactMask(T, Fl, N, sample=sample)
        mask = cm.dense()[:Fl * N, :Fl * N].contiguous()          # mask[:F*N, :F*N] of the write pass (:105-114)
        kf = int(mask[::N].sum().item())                            # sum_f K_f (one row per frame)
        flops = 4 * 64 * heads * 2 * N * kf
        q = torch.randn((2, heads, Fl * N, 64), device=dev, dtype=dtype)
        k, v = torch.randn_like(q), torch.randn_like(q)
        res = {}
        for bname, b in backends.items():
            try:
                with sdpa_kernel(b):
                    for _ in range(2):
                        F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                res[bname] = {"ms": round(ms, 4), "tflops": round(flops / ms * 1e-9, 1)}
            except Exception as e:   # noqa: BLE001  (a backend that does not take this call)
                res[bname] = {"error": str(e).splitlines()[0][:160]}
                torch.cuda.synchronize()
        out["layers"][name] = {"tokens": N, "channels": C, "heads": heads, "per_step": count,
                               "algorithmic_tflop": round(flops / 1e12, 4), "backends": res}
        ok = [r["ms"] for r in res.values() if "ms" in r]
        if ok:
            step_ms[name] = min(ok) * count
        del q, k, v, mask
        torch.cuda.empty_cache()
    if len(step_ms) == len(classes):
        total = sum(step_ms.values())
        tflop = sum(out["layers"][n]["algorithmic_tflop"] * c for n, _, _, _, c in classes)
        out["step_equivalent"] = {"attention_ms": round(total, 3), "tflops": round(tflop / total * 1e3, 1)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
