#!/usr/bin/env python
"""Per-tile timeline of the attention kernel from a -DCSA_TRACE=1 build (debug tool).

    CSA_B200_LIB=spider_b200/variants/libcsa_<trace build>.so python tools/trace_timeline.py [N C heads [F]]

Runs one write-consistent launch on the given layer class, reads the (event, SM clock) records of CTA 0 — the softmax
warp of lane quarter 0 of each Q tile and the two MMA-issue warps — and prints the median duration of every segment
of the softmax iteration and of the MMA issue loop, plus the raw events of a few steady-state tiles."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")
SLOTS, EVENTS = 4, 8192
NAMES = {1: "top", 2: "s_full", 3: "S in regs", 4: "max", 5: "o_done/rescale", 6: "token", 7: "exp issued",
         8: "p_ready", 10: "unit", 11: "epi wait", 12: "epi go", 13: "epi decoded", 14: "epi stored", 20: "qk: top", 21: "qk: k_full", 22: "qk: s_free",
         23: "pv: top", 24: "pv: v_full", 25: "pv: p_ready", 26: "pv: issued", 30: "o_done wait", 31: "o_done ok",
         **{40 + c: f"P round {c}" for c in range(8)}}


def main():
    N, C, heads = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4096, 640, 10)
    F = int(sys.argv[4]) if len(sys.argv) > 4 else 4
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
                                 b_group_rows=F * N, cb=(0, N, N))
    fn()
    torch.cuda.synchronize()
    buf = torch.zeros(SLOTS * EVENTS, dtype=torch.int64, device=dev)
    lib = native.load()
    rc = lib.csa_debug_set_trace(buf.data_ptr())
    if rc != 0:
        raise SystemExit("library was not built with -DCSA_TRACE=1: " + lib.csa_last_error().decode())
    fn()
    torch.cuda.synchronize()
    lib.csa_debug_set_trace(None)
    raw = buf.cpu().view(SLOTS, EVENTS)
    slots = []
    for sl in range(SLOTS):
        ev = [(int(x) >> 48, int(x) & ((1 << 48) - 1)) for x in raw[sl].tolist() if x != 0]
        slots.append(ev)
    t0 = min(ev[0][1] for ev in slots if ev)
    for sl, name in enumerate(("softmax Q0", "softmax Q1", "mma Q0", "mma Q1")):
        ev = slots[sl]
        print(f"== {name}: {len(ev)} events")
        seg = {}
        for (a, ta), (b, tb) in zip(ev, ev[1:]):
            seg.setdefault((a, b), []).append(tb - ta)
        for (a, b), d in sorted(seg.items(), key=lambda kv: -sum(kv[1])):
            if len(d) >= 3:
                print(f"   {str(NAMES.get(a, a)):>16s} -> {str(NAMES.get(b, b)):<16s} n={len(d):5d}  median {statistics.median(d):7.0f}"
                      f"  mean {statistics.mean(d):7.0f}  p90 {sorted(d)[int(0.9 * len(d))]:7.0f}")
        if sl < 2:
            per = [tb - ta for (a, ta), (b, tb) in zip([e for e in ev if e[0] == 6], [e for e in ev if e[0] == 6][1:])]
            if per:
                print(f"   token -> next token (tile period): median {statistics.median(per):.0f} mean {statistics.mean(per):.0f}")
    # interleaved raw view of a few steady-state tiles
    merged = sorted(((t - t0, sl, e) for sl, ev in enumerate(slots) for e, t in ev), key=lambda x: x[0])
    start = next(i for i, x in enumerate(merged) if x[0] > 60000)
    print("== raw events (clock, slot, event) from clock 60000")
    for t, sl, e in merged[start:start + 90]:
        print(f"   {t:8d}  {'  ' * sl}{('sm0', 'sm1', 'mma0', 'mma1')[sl]} {NAMES.get(e, e)}")


if __name__ == "__main__":
    main()
