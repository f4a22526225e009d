#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a B200 box; tests/test_gpu_dist.py launches it when >= 4 GPUs are
visible):   python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/gpu_dist_check.py

Every rank runs (a) the unsharded processor on the full batch — the single-GPU path that the `-m gpu` parity tests pin
to the oracle — and (b) the sharded processor on its own frames, with the fused peer-memory exchange and with the NCCL
all-gather; the sharded outputs must equal the rows of (a) within the rounding of a different key order.  Several
steps and two layer sizes, so that exchange slots are reused, buffers re-allocated and masks re-sampled."""
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spider_b200  # noqa: E402
from spider_b200 import native  # noqa: E402
from spider_b200.dist import FrameSharding  # noqa: E402
from spider_b200.install import make_processor_class  # noqa: E402


class Attn(torch.nn.Module):
    def __init__(self, c, heads):
        super().__init__()
        self.heads = heads
        self.to_q = torch.nn.Linear(c, c, bias=False)
        self.to_k = torch.nn.Linear(c, c, bias=False)
        self.to_v = torch.nn.Linear(c, c, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(c, c), torch.nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    Fl = int(sys.argv[1]) if len(sys.argv) > 1 else max(4, world // 2)
    H = W = 512                                   # 256 tokens at /32, 1024 at /16
    layers = [(256, 1280, 20), (1024, 640, 10)]
    steps = 3
    dtype = torch.bfloat16
    torch.manual_seed(0)
    attns = [Attn(c, h).to(dev, dtype) for (_, c, h) in layers]
    g = torch.Generator(device=dev).manual_seed(99)
    xs = [[torch.randn((2 * Fl, n, c), device=dev, generator=g).to(dtype) for (n, c, _) in layers] for _ in range(steps)]

    def run(exchange):
        host = spider_b200.StoryGlobals()
        host.height, host.width, host.total_count, host.id_length = H, W, len(layers), Fl
        host.write, host.cur_step, host.attn_count = True, 25, 0
        cls = make_processor_class(host)
        sh = FrameSharding(Fl, None, dev, exchange=exchange) if exchange else None
        procs = [cls(id_length=Fl, device=str(dev), dtype=torch.float16) for _ in layers]
        for p in procs:
            p.dist = sh
        random.seed(3)
        torch.manual_seed(5)
        torch.cuda.manual_seed_all(5)
        host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device=str(dev),
                                                                    dtype=torch.float16)
        if sh is not None:
            sh.sync_masks(host.mask1024, host.mask4096)
        outs = []
        real = random.random
        random.random = lambda: 0.999
        try:
            with torch.no_grad():
                for s in range(steps):
                    for li, p in enumerate(procs):
                        x = xs[s][li]
                        if sh is not None:
                            x = x[sh.cfg * Fl:(sh.cfg + 1) * Fl][sh.f0:sh.f0 + sh.frames_local].contiguous()
                        outs.append(p(attns[li], x).float())
        finally:
            random.random = real
        torch.cuda.synchronize()
        return outs, sh

    full, _ = run(None)
    worst = {}
    for exchange in ("p2p", "nccl"):
        outs, sh = run(exchange)
        err = 0.0
        for got, ref in zip(outs, full):
            want = ref[sh.cfg * Fl + sh.f0:sh.cfg * Fl + sh.f0 + sh.frames_local]
            err = max(err, (got - want).abs().max().item())
        worst[exchange] = err
        if sh.gc > 1 and exchange == "p2p":
            assert sh.peers is not None and native.LAUNCHES["csa_peer_scatter_kv"] >= steps * len(layers)
    t = torch.tensor([worst["p2p"], worst["nccl"]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"gpu_dist_check world={world} F={Fl}: max-abs vs unsharded  p2p {t[0].item():.3e}  nccl {t[1].item():.3e}",
              flush=True)
    ok = t[0].item() < 8e-3 and t[1].item() < 8e-3     # same math, different key order: bf16 output rounding
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("gpu_dist_check ok", flush=True)


if __name__ == "__main__":
    main()
