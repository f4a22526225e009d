#!/usr/bin/env python
"""How does the attention launch time depend on the number of scheduling rounds?  (debug tool)
Times csa_attn_fwd on the 32x32-class layer for head counts that give 1..8 whole rounds and partial rounds, with the
tail split on and off.  time = fixed + rounds * unit_time is the model the tail split is built on."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")


def run(F, N, heads, split, iters=20):
    C = heads * 64
    torch.manual_seed(0)
    T = F + 1
    q = torch.randn(2 * F * N, C, device=dev, dtype=torch.bfloat16)
    k = torch.randn_like(q)
    v = torch.randn_like(q)
    o = torch.empty_like(q)
    sample = torch.rand((T * N,), device=dev) < 0.5
    cm = csa_masks.CompactMask(T, F, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(dev)
    k_s, v_s, cap = native.gather_kv(k, v, F * N, 2, s_idx, s_count, F * N)
    fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=F, n_q=N, k_a=k_s, v_a=v_s,
                                 a_group_rows=cap, ranges=ranges, range_base=0, range_step=1, k_b=k, v_b=v,
                                 b_group_rows=F * N, cb=(0, N, N), split=split)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    units = 2 * F * heads * ((N + 255) // 256)
    ll = native.last_launch()
    print(f"N={N} heads={heads:3d} units={units:5d} rounds={units / 148:5.2f} split={int(split)} -> {ms * 1e3:8.1f} us  "
          f"{ms * 1e3 / (units / 148):7.1f} us/round  {ll}", flush=True)


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for heads in (4, 9, 10, 18, 19, 20, 28, 37):
        for split in (False, True):
            run(4, N, heads, split)
