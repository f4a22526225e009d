#!/usr/bin/env python
"""2-GPU probe of the peer-memory plumbing (torchrun --nproc-per-node 2): which way of mapping a peer's buffer lets
OUR kernels store into it.  Debug tool."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import native  # noqa: E402
from spider_b200 import dist as sd  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
native.ensure_device(dev)


def say(*a):
    print(f"[rank {rank}]", *a, flush=True)


def try_store(peer_buf: torch.Tensor, tag: str):
    """gather_rows kernel on THIS device writing 64 rows of 256 B into the peer's buffer."""
    src = torch.full((64, 128), float(rank + 1), dtype=torch.bfloat16, device=dev)
    idx = torch.arange(64, dtype=torch.int32, device=dev)
    try:
        out = peer_buf[:64 * 256].view(torch.bfloat16).view(64, 128)
        native.gather_rows(src, idx, 64, out=out)
        torch.cuda.synchronize(dev)
        say(tag, "store ok; ptr device reported by torch:", peer_buf.device)
        return True
    except Exception as e:   # noqa: BLE001
        say(tag, "FAILED:", repr(e)[:300])
        return False


for mode in sys.argv[1:] or ["own", "torch"]:
    local = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    handles = [None] * world
    if mode == "own":
        dist.all_gather_object(handles, native.ipc_export(local))
        peer = native.ipc_import(handles[1 - rank], dev)
    else:
        dist.all_gather_object(handles, sd._export_tensor(local))
        peer = sd._import_tensor(handles[1 - rank])
        with torch.cuda.device(dev):
            native.enable_peer_access(peer.device.index)
    ok = try_store(peer, mode)
    dist.barrier()
    torch.cuda.synchronize(dev)
    if ok:
        got = local[:64 * 256].view(torch.bfloat16).float()
        say(mode, "my buffer now holds", got.min().item(), got.max().item(), "(expect", float(2 - rank), ")")
    dist.barrier()
dist.destroy_process_group()
