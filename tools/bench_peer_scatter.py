#!/usr/bin/env python
"""Single-GPU timing of csa_peer_scatter_kv (the fused gather + exchange kernel) with every "peer" buffer in LOCAL
memory: what the kernel costs when the links are not the limit (HBM-bound: one read of the sampled rows, n_peers
writes).  The multi-GPU numbers are in profiles/*_scale; this is the kernel's own roofline line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")


def run(F, N, C, n_peers, fr, iters=20):
    torch.manual_seed(0)
    T = F + 1
    k = torch.randn(fr * N, C, device=dev, dtype=torch.bfloat16)
    v = torch.randn_like(k)
    sample = torch.rand((T * N,), device=dev) < 0.5
    cm = csa_masks.CompactMask(T, F, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(dev)
    rows = F * N + native.CSA_TILE
    kd = [torch.zeros(rows, C, device=dev, dtype=torch.bfloat16) for _ in range(n_peers)]
    vd = [torch.zeros(rows, C, device=dev, dtype=torch.bfloat16) for _ in range(n_peers)]
    flags = [torch.zeros((3, 8), dtype=torch.int32, device=dev) for _ in range(n_peers)]
    me = 1
    epoch = 0

    def fn():
        nonlocal epoch
        epoch += 1
        native.peer_scatter_kv(k, v, s_idx, fr * N, 0, kd, vd, [f[0] for f in flags], me, epoch, flags[me][1], 0,
                               flags[me][2], ranges=ranges, frames_per_peer=fr, idx_adjust=-me * fr * N)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # the 20 launches are issued from ONE call into the library (csa_run_batch): the Python wrapper costs ~28 us per
    # call, more than the kernel takes on the small layers
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    native.begin_batch(k)
    for _ in range(iters):
        fn()
    e0.record()
    native.flush_batch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rh = ranges.cpu().tolist()
    lo, hi = rh[me * fr][1], rh[(me + 1) * fr - 1][2]
    cnt = hi - lo
    # check: the rows landed where the S order says, on every "peer"
    want = k[(s_idx[lo:hi].long() - me * fr * N)]
    ok = all(torch.equal(kd[r][lo:hi], want) for r in range(n_peers)) and int(flags[0][0][me]) == epoch
    byts = 2 * cnt * C * 2 * (1 + n_peers)
    return {"F": F, "N": N, "C": C, "peers": n_peers, "frames_local": fr, "rows": cnt, "ms": round(ms, 5),
            "GBps": round(byts / ms * 1e-6, 1), "bytes": byts, "ok": ok}


if __name__ == "__main__":
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6550.0)
    for (F, N, C, n, fr) in [(16, 4096, 640, 4, 4), (16, 1024, 1280, 4, 4), (4, 4096, 640, 4, 1), (4, 1024, 1280, 4, 1)]:
        r = run(F, N, C, n, fr)
        r["frac_of_hbm_peak"] = round(r["GBps"] / hbm, 3)
        print(json.dumps(r), flush=True)
