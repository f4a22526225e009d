"""Multi-GPU Consistent Self-Attention: (CFG half, frame) units sharded over the GPUs of one box.

The reference is single-GPU; this is the B200 scale-out of its write pass.  The two CFG halves never
attend each other (StoryDiffusion/Comic_Generation.py:148 folds the frames of ONE half into a sequence), and inside a
half frame f attends ``S u block_f`` where S — the sampled key positions — is one list shared by all frames
(StoryDiffusion/utils/gradio_utils.py:257-278).  So with G ranks (G even):

  * ranks ``[0, G/2)`` hold the unconditional half, ranks ``[G/2, G)`` the conditional half; inside a half the F
    identity frames are split into contiguous runs of ``F / (G/2)`` frames per rank;
  * every layer runs on the local frames only; in the consistent branch each rank gathers the sampled rows of ITS
    frames from its K/V projections (``csa_gather_rows``, HBM-bound), the slabs are all-gathered inside the half over
    NVLink (NCCL; nothing crosses between the halves), compacted into the same ``K[S], V[S]`` buffers the single-GPU
    path builds (``csa_gather_kv``), and the local frames attend ``two runs of that buffer + their own block`` with
    the unchanged attention kernel;
  * G == 2 needs no exchange at all (one half per GPU).

All sizes are known on every rank because S is global: the sample vector is broadcast from rank 0 whenever it is
re-drawn (``sync_masks``), and the per-rank run lengths are read back once per step and resolution (one host sync
per mask, shared by all layers).  The branch gate uses Python's ``random`` and must be seeded identically on every
rank (as ``setup_seed`` does, Comic_Generation.py:35-40); ``check_lockstep`` asserts it.

Read passes (one generated frame per call, batch 2) are independent given the bank and are not sharded here: run
them on the rank(s) that hold the bank.  A sharded write pass keeps a sharded bank.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import native


class ShardPlan:
    """Per (mask, sharding): who sends how many sampled rows, and where they land in the compact buffer."""

    def __init__(self, cm, sh: "FrameSharding", device):
        F, N = cm.id_length, cm.n_tokens
        s_idx, s_count, ranges = cm.sample_list(device)
        rh = ranges.cpu().tolist()                     # the one host sync per mask
        self.s_idx, self.s_count, self.ranges = s_idx, s_count, ranges
        self.total = int(rh[F][1])
        fr = sh.frames_local
        # run of S that falls into the frames of rank r of this half: [lo_r, hi_r)
        self.lo = [rh[r * fr][1] for r in range(sh.gc)]
        self.hi = [rh[r * fr + fr - 1][2] for r in range(sh.gc)]
        self.counts = [h - l for l, h in zip(self.lo, self.hi)]
        self.pad = max(8, (max(self.counts) + 7) // 8 * 8)
        me = sh.rank_in_half
        self.count_me = self.counts[me]
        # positions of my sampled rows inside my local K/V (frames f0 .. f0+fr-1 are rows [0, fr*N))
        self.local_idx = (s_idx[self.lo[me]:self.hi[me]] - sh.f0 * N).contiguous()
        # compact position i of S -> row of the all-gathered slab matrix [(rank, K|V, pad rows), C]
        parts = [torch.arange(c, dtype=torch.int32, device=device) + 2 * r * self.pad
                 for r, c in enumerate(self.counts)]
        self.slab_map = torch.cat(parts) if parts else torch.zeros((0,), dtype=torch.int32, device=device)


class FrameSharding:
    """Static assignment of (CFG half, frame) units to the ranks of ``group`` (default: the world)."""

    def __init__(self, id_length: int, group=None, device: Optional[torch.device] = None):
        if not dist.is_initialized():
            raise RuntimeError("FrameSharding needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        if self.world < 2 or self.world % 2:
            raise ValueError(f"frame sharding needs an even number of ranks >= 2 (one CFG half per rank group), "
                             f"got {self.world}")
        self.gc = self.world // 2                 # ranks per CFG half
        if id_length % self.gc:
            raise ValueError(f"id_length {id_length} is not divisible by the {self.gc} ranks of a CFG half")
        self.id_length = id_length
        self.frames_local = id_length // self.gc
        self.cfg = self.rank // self.gc           # 0: unconditional half, 1: conditional half
        self.rank_in_half = self.rank % self.gc
        self.f0 = self.rank_in_half * self.frames_local
        self.local_batch = self.frames_local      # latents per rank (all of one CFG half)
        self.half_group = None
        if self.gc > 1:
            # every rank has to take part in the creation of both groups
            world_ranks = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
            for c in range(2):
                g = dist.new_group(ranks=[world_ranks[c * self.gc + i] for i in range(self.gc)])
                if c == self.cfg:
                    self.half_group = g
        self.bytes_exchanged = 0                  # received bytes, for reporting

    # ---------------------------------------------------------------------------------------------- lock-step
    def sync_masks(self, *masks) -> None:
        """Broadcast freshly sampled vectors from rank 0 so that every rank compacts the same S."""
        for cm in masks:
            sample = getattr(cm, "_sample", None)
            if sample is None:
                raise ValueError("sharded runs need masks in compact (sample-vector) form")
            if self.device is not None and sample.device != torch.device(self.device):
                sample = cm._sample = sample.to(self.device)
            buf = sample.view(torch.uint8)
            dist.broadcast(buf, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)

    def check_lockstep(self, value: float) -> None:
        """Debug aid: all ranks must have drawn the same gate value (same Python ``random`` seed)."""
        t = torch.tensor([value, -value], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        if float(t[0]) != value or float(-t[1]) != value:
            raise RuntimeError("ranks diverged: the branch gate drew different random numbers (seed Python's "
                               "`random` identically on every rank)")

    # ---------------------------------------------------------------------------------------------- the layer
    def plan(self, cm, device) -> ShardPlan:
        p = getattr(cm, "_shard_plan", None)
        if p is None or p[0] is not self:
            p = (self, ShardPlan(cm, self, device))
            cm._shard_plan = p
        return p[1]

    def attn_write(self, q, k, v, o, N, heads, cm, Fl):
        """Write-mode consistent attention of the local frames (``__call1__``, Comic_Generation.py:129-196 with
        mask[:F*N,:F*N]) with the sampled rows of the other ranks' frames fetched over NVLink."""
        if Fl != self.id_length:
            raise ValueError(f"sharding was built for id_length {self.id_length}, processor has {Fl}")
        fr = self.frames_local
        if q.shape[0] != fr * N:
            raise ValueError(f"rank {self.rank} expects {fr} local frames ({fr * N} rows), got {q.shape[0]} rows")
        if not cm.shared_sample:
            raise ValueError("sharded consistent attention needs masks whose rows share one sample vector")
        C = q.shape[1]
        pl = self.plan(cm, q.device)
        if self.gc == 1:
            # one CFG half per GPU: everything is local, same as the single-GPU path with one group
            k_s, v_s, cap = native.gather_kv(k, v, Fl * N, 1, pl.s_idx, pl.s_count, Fl * N)
        else:
            send = torch.empty((2, pl.pad, C), dtype=k.dtype, device=k.device)
            if pl.count_me > 0:
                native.gather_rows(k, pl.local_idx, pl.count_me, out=send[0])
                native.gather_rows(v, pl.local_idx, pl.count_me, out=send[1])
            recv = torch.empty((self.gc, 2, pl.pad, C), dtype=k.dtype, device=k.device)
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.half_group)
            self.bytes_exchanged += (self.gc - 1) * send.numel() * send.element_size()
            flat = recv.view(-1, C)
            # slab rows -> the compact K[S], V[S] buffers of the single-GPU path (zero tail included)
            k_s, v_s, cap = native.gather_kv(flat[:-pl.pad], flat[pl.pad:], flat.shape[0] - pl.pad, 1, pl.slab_map,
                                             pl.s_count, Fl * N)
        native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=fr, n_q=N,
                        k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=pl.ranges, range_base=self.f0, range_step=1,
                        k_b=k, v_b=v, b_group_rows=fr * N, cb=(0, N, N))
        return o
