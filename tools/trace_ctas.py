#!/usr/bin/env python
"""Per-CTA unit boundaries of one attention launch from a -DCSA_TRACE=1 build (debug tool): when does every CTA finish
each of its work units (whole units and split pieces), relative to the first CTA's start?  Answers what the last,
partial scheduling round really costs.

    CSA_B200_LIB=spider_b200/variants/libcsa_<trace build>.so python tools/trace_ctas.py [N C heads [F]]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")
SLOTS, EVENTS, MARKS = 4, 8192, 64


def main():
    N, C, heads = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 1280, 20)
    F = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    torch.manual_seed(0)
    T = F + 1
    q = torch.randn(2 * F * N, C, device=dev, dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    o = torch.empty_like(q)
    sample = torch.rand((T * N,), device=dev) < 0.5
    cm = csa_masks.CompactMask(T, F, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(dev)
    k_s, v_s, cap = native.gather_kv(k, v, F * N, 2, s_idx, s_count, F * N)
    fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=F, n_q=N, k_a=k_s, v_a=v_s,
                                 a_group_rows=cap, ranges=ranges, range_base=0, range_step=1, k_b=k, v_b=v,
                                 b_group_rows=F * N, cb=(0, N, N), split=os.environ.get('CSA_NO_SPLIT') != '1')
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf = torch.zeros(SLOTS * EVENTS + 148 * MARKS, dtype=torch.int64, device=dev)
    lib = native.load()
    if lib.csa_debug_set_trace(buf.data_ptr()) != 0:
        raise SystemExit("library was not built with -DCSA_TRACE=1: " + lib.csa_last_error().decode())
    fn()
    torch.cuda.synchronize()
    lib.csa_debug_set_trace(None)
    print("launch:", native.last_launch())
    raw = buf.cpu()[SLOTS * EVENTS:].view(148, MARKS).tolist()
    ctas = []
    for row in raw:
        ev = [((x >> 56) & 0xff, x & ((1 << 56) - 1)) for x in row if x != 0]
        if ev:
            ctas.append(ev)
    t0 = min(ev[0][1] for ev in ctas)
    starts = [ev[0][1] - t0 for ev in ctas]
    ends = [ev[-1][1] - t0 for ev in ctas]
    print(f"{len(ctas)} CTAs; start spread {max(starts) / 1e3:.1f} us; last unit ends: min {min(ends) / 1e3:.1f} "
          f"median {statistics.median(ends) / 1e3:.1f} max {max(ends) / 1e3:.1f} us")
    depth = max(len(ev) for ev in ctas)
    for kth in range(1, depth):
        durs, tags, at = [], set(), []
        for ev in ctas:
            if len(ev) > kth:
                durs.append(ev[kth][1] - ev[kth - 1][1])
                tags.add(ev[kth][0])
                at.append(ev[kth][1] - t0)
        kind = "/".join({2: "whole", 3: "piece"}.get(t, str(t)) for t in sorted(tags))
        print(f"  unit {kth:2d} ({kind:11s}) on {len(durs):3d} CTAs: duration min {min(durs) / 1e3:6.1f} median "
              f"{statistics.median(durs) / 1e3:6.1f} max {max(durs) / 1e3:6.1f} us; finished at median "
              f"{statistics.median(at) / 1e3:6.1f} max {max(at) / 1e3:6.1f} us")


if __name__ == "__main__":
    main()
