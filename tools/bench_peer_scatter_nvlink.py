#!/usr/bin/env python
"""csa_peer_scatter_kv across REAL NVLink peers, from ONE process (so that ncu may wrap it: link counters
nvltx__bytes / nvlrx__bytes).  GPU 0 runs the kernel; the K[S] / V[S] buffers and flag words of the other "ranks" live on
GPUs 1 .. P-1 of the box (peer access enabled from GPU 0), rank `me`'s own buffers on GPU 0.  Checks that the rows and
the arrival flags landed on every GPU, then times the kernel with CUDA events: bytes over the links =
2 (K, V) x rows x C x 2 B x (P - 1) per launch, against the measured 770 GB/s per direction of a B200 (B200_PROFILING.md).

    gpurun --gpus 4 -- 'python tools/bench_peer_scatter_nvlink.py; ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,\\
        gpu__time_duration.sum -k regex:peer_scatter -c 8 python tools/bench_peer_scatter_nvlink.py --iters 2'"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402


def run(F, N, C, peers, fr, iters):
    dev0 = torch.device("cuda:0")
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    T = F + 1
    me = 0
    k = torch.randn(fr * N, C, device=dev0, dtype=torch.bfloat16)
    v = torch.randn_like(k)
    sample = torch.rand((T * N,), device=dev0) < 0.5
    cm = csa_masks.CompactMask(T, F, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(dev0)
    rows = F * N + native.CSA_TILE
    devs = [torch.device("cuda", r) for r in range(peers)]
    for r in range(1, peers):
        native.enable_peer_access(r)            # GPU 0 -> GPU r
    kd = [torch.zeros(rows, C, device=d, dtype=torch.bfloat16) for d in devs]
    vd = [torch.zeros(rows, C, device=d, dtype=torch.bfloat16) for d in devs]
    flags = [torch.zeros((3, 8), dtype=torch.int32, device=d) for d in devs]
    for d in devs:
        torch.cuda.synchronize(d)
    epoch = 0

    def fn():
        nonlocal epoch
        epoch += 1
        native.peer_scatter_kv(k, v, s_idx, fr * N, 0, kd, vd, [f[0] for f in flags], me, epoch, flags[me][1], 0,
                               flags[me][2], ranges=ranges, frames_per_peer=fr, idx_adjust=-me * fr * N)
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    native.begin_batch(k)
    for _ in range(iters):
        fn()
    e0.record()
    native.flush_batch()
    e1.record()
    torch.cuda.synchronize(dev0)
    ms = e0.elapsed_time(e1) / iters
    rh = ranges.cpu().tolist()
    lo, hi = rh[me * fr][1], rh[(me + 1) * fr - 1][2]
    cnt = hi - lo
    want_k = k[(s_idx[lo:hi].long() - me * fr * N)]
    want_v = v[(s_idx[lo:hi].long() - me * fr * N)]
    ok = True
    for r in range(peers):
        torch.cuda.synchronize(devs[r])
        ok = ok and torch.equal(kd[r][lo:hi].to(dev0), want_k) and torch.equal(vd[r][lo:hi].to(dev0), want_v)
        ok = ok and int(flags[r][0][me]) == epoch
    link_bytes = 2 * cnt * C * 2 * (peers - 1)
    return {"F": F, "N": N, "C": C, "gpus": peers, "frames_local": fr, "rows": cnt, "ms": round(ms, 5),
            "link_bytes_per_launch": link_bytes, "link_GBps_tx": round(link_bytes / ms * 1e-6, 1),
            "frac_of_770_GBps": round(link_bytes / ms * 1e-6 / 770.0, 3), "ok": bool(ok)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    n = torch.cuda.device_count()
    if n < 2:
        raise SystemExit("needs >= 2 GPUs on one box")
    peers = 4 if n >= 4 else 2
    # (F, N, C, frames of this rank): config 4 shapes (16 frames over the ranks of a CFG half) and the 4-frame story
    for (F, N, C, fr) in [(16, 4096, 640, 16 // peers), (16, 1024, 1280, 16 // peers), (4, 4096, 640, max(1, 4 // peers)),
                          (4, 1024, 1280, max(1, 4 // peers))]:
        print(json.dumps(run(F, N, C, peers, fr, a.iters)), flush=True)
