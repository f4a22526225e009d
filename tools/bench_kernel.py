#!/usr/bin/env python
"""Kernel-only timing of csa_attn_fwd (write-consistent, pre-gathered sample) on the two SDXL layer classes.
Used for tuning sweeps: `CSA_B200_LIB=spider_b200/variants/libcsa_x.so python tools/bench_kernel.py [tag]`.
Prints one line per layer class: ms, algorithmic TFLOP/s, max-abs error vs torch SDPA on one (group, frame)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("CSA_B200_LIB", "default"))


def run(F, N, C, heads, sa=0.5, dtype=torch.bfloat16, iters=20):
    torch.manual_seed(0)
    T = F + 1
    q = torch.randn(2 * F * N, C, device=dev, dtype=dtype)
    k = torch.randn(2 * F * N, C, device=dev, dtype=dtype)
    v = torch.randn(2 * F * N, C, device=dev, dtype=dtype)
    o = torch.empty_like(q)
    sample = torch.rand((T * N,), device=dev) < sa
    cm = csa_masks.CompactMask(T, F, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(dev)
    k_s, v_s, cap = native.gather_kv(k, v, F * N, 2, s_idx, s_count, F * N)
    fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=F, n_q=N, k_a=k_s, v_a=v_s,
                                 a_group_rows=cap, ranges=ranges, range_base=0, range_step=1, k_b=k, v_b=v,
                                 b_group_rows=F * N, cb=(0, N, N), split=os.environ.get('CSA_NO_SPLIT') != '1')
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # flush L2 between iterations is not needed: q/k/v/o of one launch (>= 84 MB) + rotation over 4 buffers
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    idx, counts = cm.lists(dev)
    kf = counts[:F].sum().item()
    flops = 4 * 64 * heads * 2 * N * kf
    # spot check: group 1, last frame
    g, f = 1, F - 1
    keys = idx[f, :counts[f].item()].long() + g * F * N
    qs = slice((g * F + f) * N, (g * F + f + 1) * N)
    qh = q[qs].float().view(N, heads, 64).transpose(0, 1)[None]
    kh = k[keys].float().view(-1, heads, 64).transpose(0, 1)[None]
    vh = v[keys].float().view(-1, heads, 64).transpose(0, 1)[None]
    ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)[0].transpose(0, 1).reshape(N, C)
    err = (o[qs].float() - ref).abs().max().item()
    print(f"{tag:<22s} F={F} N={N:5d} C={C:4d}: {ms:7.4f} ms {flops / ms * 1e-9:7.1f} TFLOP/s  max-abs {err:.2e}"
          f"  {native.last_launch()}", flush=True)
    return ms


if __name__ == "__main__":
    t64 = run(4, 4096, 640, 10)
    t32 = run(4, 1024, 1280, 20)
    print(f"{tag:<22s} step-equivalent attention time (6 x 64x64 + 30 x 32x32): {6 * t64 + 30 * t32:.3f} ms "
          f"-> {8.4196 / (6 * t64 + 30 * t32) * 1e3:.1f} TFLOP/s", flush=True)
