#!/usr/bin/env python
"""Multi-GPU parity + stress check (run under torchrun on a B200 box; tests/test_gpu_dist.py and smoke() launch it
when enough GPUs are visible):

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/gpu_dist_check.py \
        [--frames F] [--steps S] [--skew] [--graph] [--no-read]

Every rank runs (a) the unsharded processor on the full batch — the single-GPU path that the `-m gpu` parity tests pin
to the oracle — and (b) the sharded processor on its own frames, with the fused peer-memory exchange (eagerly and, with
--graph, as ONE captured CUDA graph per step replayed S times) and with the NCCL all-gather; the sharded outputs must
equal the rows of (a) within the rounding of a different key order.  S steps x two layer sizes = 2*S exchange epochs:
slots are reused, epochs advance in device memory, masks are re-sampled in place.  --skew makes one rank of every CFG
half late by a spinning kernel before random layer calls (the flag protocol must hold with ranks a layer apart).
Afterwards (unless --no-read) the story FINISHES: every rank all-gathers the sharded id_bank and generates its own
frame through the read pass, checked against the unsharded processor reading the unsharded bank."""
import argparse
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spider_b200  # noqa: E402
from spider_b200 import native  # noqa: E402
from spider_b200.dist import FrameSharding  # noqa: E402
from spider_b200.graph import StepGraph  # noqa: E402
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skew", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--no-read", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    Fl = args.frames or max(4, world // 2)
    H = W = 512                                   # 256 tokens at /32, 1024 at /16
    layers = [(256, 1280, 20), (1024, 640, 10)]
    steps = args.steps
    dtype = torch.bfloat16
    torch.manual_seed(0)
    attns = [Attn(c, h).to(dev, dtype) for (_, c, h) in layers]
    xbuf = [torch.empty((2 * Fl, n, c), device=dev, dtype=dtype) for (n, c, _) in layers]   # refreshed every step

    def fill_inputs(step):
        g = torch.Generator(device=dev).manual_seed(1000 + step)
        for b in xbuf:
            b.copy_(torch.randn(b.shape, device=dev, generator=g).to(dtype))

    def run(exchange, graph=False, skew=False):
        host = spider_b200.StoryGlobals()
        host.height, host.width, host.total_count, host.id_length = H, W, len(layers), Fl
        host.write, host.cur_step, host.attn_count = True, 25, 0
        cls = make_processor_class(host)
        sh = FrameSharding(Fl, None, dev, exchange=exchange) if exchange else None
        if sh is not None and graph:
            sh.mask_sync = "seeded"               # a captured step cannot broadcast; every rank seeds alike
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
        local_x = xbuf
        if sh is not None:
            rows = slice(sh.cfg * Fl + sh.f0, sh.cfg * Fl + sh.f0 + sh.frames_local)
            local_x = [torch.empty_like(b[rows]) for b in xbuf]
        late = sh is not None and skew and sh.rank_in_half == sh.gc - 1
        rnd = random.Random(7 + rank)

        def step_fn():
            host.cur_step, host.attn_count = 25, 0      # the bank entry of step 25 is overwritten (bounded memory)
            outs = []
            for li, p in enumerate(procs):
                if late and not graph and rnd.random() < 0.3:
                    torch.cuda._sleep(int(2e6))           # ~1 ms: this rank falls a layer or more behind its peers
                outs.append(p(attns[li], local_x[li]))
            return outs

        outs = []
        real = random.random
        random.random = lambda: 0.999
        try:
            with torch.no_grad():
                sg = None
                for s in range(steps):
                    fill_inputs(s)
                    if sh is not None:
                        for b, lx in zip(xbuf, local_x):
                            lx.copy_(b[rows])
                    if graph:
                        if sg is None:
                            # the capture's warm-up is step 0 itself (inputs in place); replays follow
                            sg = StepGraph(step_fn, dev, warmup=0)
                            o = [t.clone() for t in step_fn()]
                            sg.capture()
                        else:
                            if late and rnd.random() < 0.3:
                                torch.cuda._sleep(int(2e6))
                            o = [t.clone() for t in sg.replay()]
                    else:
                        o = [t.clone() for t in step_fn()]
                    outs.append([t.float() for t in o])
        finally:
            random.random = real
        torch.cuda.synchronize()
        return outs, sh, host, procs

    full, _, host_full, procs_full = run(None)
    worst = {}
    modes = [("p2p", False), ("nccl", False)] + ([("p2p", True)] if args.graph else [])
    kept = None
    for exchange, graph in modes:
        scatter0 = native.LAUNCHES["csa_peer_scatter_kv"]
        outs, sh, host_s, procs_s = run(exchange, graph=graph, skew=args.skew)
        err = 0.0
        for got_step, ref_step in zip(outs, full):
            for got, ref in zip(got_step, ref_step):
                want = ref[sh.cfg * Fl + sh.f0:sh.cfg * Fl + sh.f0 + sh.frames_local]
                err = max(err, (got - want).abs().max().item())
        worst[(exchange, graph)] = err
        if sh.gc > 1 and exchange == "p2p" and not graph:
            assert sh.peers is not None
            scatters = native.LAUNCHES["csa_peer_scatter_kv"] - scatter0
            # fused exchange: the K|V projection delivers the rows itself — no scatter launch at all
            assert scatters == (0 if sh.fused_exchange else steps * len(layers)), scatters
        if exchange == "p2p" and not graph:
            kept = (sh, host_s, procs_s)
    # ------------------------------------------------------------------ the story finishes: frame-parallel reads
    read_err = 0.0
    if not args.no_read:
        sh, host_s, procs_s = kept
        gx = torch.Generator(device=dev).manual_seed(4242 + rank)
        xr = [torch.randn((2, n, c), device=dev, generator=gx).to(dtype) for (n, c, _) in layers]
        real = random.random
        random.random = lambda: 0.999
        try:
            with torch.no_grad():
                for host_x, procs_x in ((host_full, procs_full), (host_s, procs_s)):
                    host_x.write, host_x.cur_step, host_x.attn_count = False, 25, 0
                # both runs ended on the same generator stream position: same masks for the read step
                torch.cuda.manual_seed_all(77)
                host_full.mask1024.resample_(0.5)
                host_full.mask4096.resample_(0.5)
                torch.cuda.manual_seed_all(77)
                host_s.mask1024.resample_(0.5)
                host_s.mask4096.resample_(0.5)
                want = [p(attns[li], xr[li]).float() for li, p in enumerate(procs_full)]
                got = [p(attns[li], xr[li]).float() for li, p in enumerate(procs_s)]
                for li, p in enumerate(procs_s):
                    assert p.id_bank[25].k.shape[0] == 2 * Fl * layers[li][0], "bank entry was not made whole"
        finally:
            random.random = real
        torch.cuda.synchronize()
        read_err = max((g - w).abs().max().item() for g, w in zip(got, want))
    vals = [worst[m] for m in modes] + [read_err]
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        names = [f"{e}{'+graph' if g else ''}" for e, g in modes] + ["read-after-gather"]
        print(f"gpu_dist_check world={world} F={Fl} steps={steps} ({2 * steps} exchange epochs) skew={args.skew}: "
              "max-abs vs unsharded  " + "  ".join(f"{n} {v:.3e}" for n, v in zip(names, t.tolist())), flush=True)
    ok = all(v < 8e-3 for v in t.tolist())     # same math, different key order: bf16 output rounding
    assert native.debug_stuck() is None
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("gpu_dist_check ok", flush=True)


if __name__ == "__main__":
    main()
