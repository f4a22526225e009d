#!/usr/bin/env python
"""BASELINE config 5 (roofline / scaling sweep) in ONE process per rank: sa32 = sa64 x frames x resolution, every point
timed like bench.py times its workload (warm-up, one captured CUDA graph per denoise step, K timed replays, CUDA
events, max over ranks).  Prints one JSON line per point and a markdown table.

    python tools/sweep_inproc.py [--sas 0 0.25 0.5 1] [--frames 2 4 8 16] [--res 768 1024 1536] [--placement up|all]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_inproc.py --gpus 8 ..."""
import argparse
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--sas", type=float, nargs="+", default=[0.0, 0.25, 0.5, 1.0])
    ap.add_argument("--frames", type=int, nargs="+", default=[2, 4, 8, 16])
    ap.add_argument("--res", type=int, nargs="+", default=[768, 1024, 1536])
    ap.add_argument("--placement", choices=["up", "all"], default="up")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", choices=["bf16", "fp16"], default="bf16")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch.distributed as dist
    from spider_b200 import native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    native.ensure_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[a.dtype]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = bench.load_peaks()
    rows = []
    shard_cache = {}                 # one FrameSharding (process groups, mapped peer buffers) per frame count
    random.random = lambda: 0.999    # gate forced open, as in bench.py
    for res in a.res:
        for F in a.frames:
            if world > 2 and F % (world // 2):
                continue
            for sa in a.sas:
                args = argparse.Namespace(res=res, sa=sa, placement=a.placement, module_projections=False,
                                          exchange=a.exchange, share_weights=True, graph=a.graph, frames=F,
                                          projections="own", no_fused_gather=False, no_fused_qkv=False)
                try:
                    wl = bench.Workload(args, F, dev, dtype, world, rank, shard_cache)
                    with torch.no_grad():
                        wl.step(count=False)
                    m = bench.time_workload(wl, args, a.steps, a.warmup, barrier, rank)
                    flops = wl.flops()
                    vals = [m["ms_total"], m["attn_ms"]]
                    if world > 1:
                        t = torch.tensor(vals, device=dev, dtype=torch.float64)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        vals = t.tolist()
                    ms_step = vals[0] / a.steps
                    kern = flops / a.steps * m["attn_steps"] / world / (vals[1] * 1e-3) / 1e12 if vals[1] > 0 else 0.0
                    row = {"res": res, "frames": F, "sa": sa, "n_gpus": world, "placement": a.placement,
                           "layers": len(wl.plan), "ms_per_step": round(ms_step, 4),
                           "tflop_per_step": round(flops / a.steps / 1e12, 3),
                           "value_tflops": round(flops / a.steps / (ms_step * 1e-3) / 1e12, 1),
                           "attn_kernel_tflops_per_gpu": round(kern, 1),
                           "attn_frac_of_peak": round(kern / peaks["bf16_tflops"], 4), "issue": m["mode"]}
                    del m, wl
                except Exception as e:   # noqa: BLE001 - a point that does not fit is reported, the sweep goes on
                    row = {"res": res, "frames": F, "sa": sa, "n_gpus": world, "error": f"{type(e).__name__}: {e}"[:200]}
                    native.abort_batch()
                torch.cuda.empty_cache()
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
    if rank == 0:
        lines = ["| res | frames | sa | GPUs | ms/step | TFLOP/step | step TFLOP/s | attention kernel TFLOP/s per GPU | frac of "
                 f"{peaks['bf16_tflops']:.0f} |", "|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            if "error" in r:
                lines.append(f"| {r['res']} | {r['frames']} | {r['sa']} | {r['n_gpus']} | {r['error']} | | | | |")
            else:
                lines.append(f"| {r['res']} | {r['frames']} | {r['sa']} | {r['n_gpus']} | {r['ms_per_step']} | "
                             f"{r['tflop_per_step']} | {r['value_tflops']} | {r['attn_kernel_tflops_per_gpu']} | "
                             f"{r['attn_frac_of_peak']} |")
        table = "\n".join(lines)
        print(table, flush=True)
        if a.out:
            with open(a.out, "w") as f:
                f.write(table + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
