"""Multi-GPU Consistent Self-Attention: (CFG half, frame) units sharded over the GPUs of one box.

The reference is single-GPU; this is the B200 scale-out of its write pass.  The two CFG halves never
attend each other (StoryDiffusion/Comic_Generation.py:148 folds the frames of ONE half into a sequence), and inside a
half frame f attends ``S u block_f`` where S — the sampled key positions — is one list shared by all frames
(StoryDiffusion/utils/gradio_utils.py:257-278).  So with G ranks (G even):

  * ranks ``[0, G/2)`` hold the unconditional half, ranks ``[G/2, G)`` the conditional half; inside a half the F
    identity frames are split into contiguous runs of ``F / (G/2)`` frames per rank;
  * every layer runs on the local frames only; in the consistent branch the sampled rows of the frames of one half
    are exchanged inside that half over NVLink (nothing crosses between the halves) into the same S-ordered
    ``K[S], V[S]`` buffers the single-GPU path builds, and the local frames attend ``two runs of that buffer + their
    own block`` with the same attention kernel.  Two exchange paths:
      - ``exchange="p2p"`` (default): ONE kernel per rank gathers its sampled rows and stores them straight into
        every peer's buffer over peer memory (``csa_peer_scatter_kv``), then raises a flag there; the receiver's
        attention kernel starts with each frame's own (local) block and waits for a peer's flag only when it reaches
        that peer's rows — the transfer is hidden behind the attention of the local keys (``PeerExchange``);
      - ``exchange="nccl"``: gather (``csa_gather_rows``) -> NCCL all-gather -> compaction (``csa_gather_kv``),
        kept as the library baseline the fused path is measured against;
  * G == 2 needs no exchange at all (one half per GPU).

All sizes are known on every rank because S is global: the sample vector is broadcast from rank 0 whenever it is
re-drawn (``sync_masks``), and the per-rank run lengths are read back once per step and resolution (one host sync
per mask, shared by all layers).  The branch gate uses Python's ``random`` and must be seeded identically on every
rank (as ``setup_seed`` does, Comic_Generation.py:35-40); ``check_lockstep`` asserts it.

Read passes (one generated frame per call, batch 2; Comic_Generation.py:441-448) are independent given the bank: a
sharded write pass leaves every rank with the K/V of its own frames, the first read of a (layer, step) entry
all-gathers it into the full bank on every rank (``gather_bank``, one NCCL collective per entry), and from then on
any rank can generate any frame — run different frames on different ranks, no per-layer exchange.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

import contextlib
import os

from . import native

_NULL = contextlib.nullcontext()


def _export_tensor(t: torch.Tensor):
    """Picklable handle of a tensor's memory for the other ranks of this box.  Device memory: a CUDA IPC handle of
    the allocation (``csa_ipc_export``).  Host memory (the gloo tests): torch's shared-memory reduction."""
    if t.is_cuda:
        return ("cuda", native.ipc_export(t))
    from multiprocessing.reduction import ForkingPickler

    import torch.multiprocessing.reductions  # noqa: F401  (registers the tensor / storage reducers)
    return ("host", bytes(ForkingPickler.dumps(t)))


def _import_tensor(handle, device) -> torch.Tensor:
    """Map a peer's buffer; nothing is copied.  Device memory is opened with THIS rank's device current, so that the
    mapping — and the peer access CUDA enables lazily with it — belongs to the device whose kernels store through it
    (a mapping opened under the exporter's device index, as torch's own CUDA IPC rebuild does, faults when written
    from another device: tools/ipc_probe.py).  Returned as a flat uint8 tensor."""
    kind, h = handle
    if kind == "cuda":
        return native.ipc_import(h, device)
    import pickle
    return pickle.loads(h).view(torch.uint8).view(-1)


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank of a CFG half when any of them cannot map a peer's buffers (no P2P path, CUDA IPC not
    permitted in this container, ...)."""


class PeerExchange:
    """Symmetric exchange buffers of one CFG half: on every rank two slots of (K[S], V[S]) byte buffers plus arrival
    (``ready``) and release (``done``) flags, each mapped into every peer (CUDA IPC + P2P access).  Layer calls are
    numbered by a monotonically increasing epoch (all ranks make the same calls in the same order); epoch e uses
    slot e % 2, so a rank may run one layer ahead of the slowest peer before ``csa_peer_scatter_kv`` has to wait
    for that peer's ``done`` flag.

    Epochs live in device memory: a call passes its number WITHIN the current step, the kernels add the
    device-resident ``epoch_base``, and ``end_step()`` — called by the processor at the step roll-over — advances the
    base by the (even) number of calls of the step with one single-thread kernel.  A denoise step captured in a CUDA
    graph therefore publishes and awaits fresh epochs at every replay (``spider_b200/graph.py``)."""

    SLOTS = 2

    def __init__(self, sh: "FrameSharding", device):
        self.sh = sh
        self.device = torch.device(device)
        self.cap_bytes = 0
        self.epoch = 0        # exchange calls since the last end_step() (the device holds the base)
        self.epoch_base = None
        self.bufs = None      # per rank: uint8 tensor [SLOTS * 2 * cap_bytes]
        self.flags = None     # per rank: int32 tensor [3, CSA_MAX_PEERS]: ready, done, {counter, ...}
        self._keep = None
        self._views = {}
        self.allocations = 0

    def ensure(self, need_bytes: int) -> None:
        """Collective over the half group: (re)allocate and re-map the buffers when a layer needs more room."""
        if need_bytes <= self.cap_bytes:
            return
        sh = self.sh
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        dist.barrier(group=sh.half_group)          # nobody is still reading or writing the old buffers
        cap = (need_bytes + 4095) // 4096 * 4096
        local = torch.zeros(self.SLOTS * 2 * cap, dtype=torch.uint8, device=self.device)
        flags = torch.zeros(3 * native.CSA_MAX_PEERS * 4, dtype=torch.uint8, device=self.device)
        if self.epoch_base is None:
            self.epoch_base = torch.zeros(1, dtype=torch.int32, device=self.device)
        else:
            self.epoch_base.zero_()                 # fresh flags: epochs restart
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)    # the zeros are in memory before anybody maps the buffers
        handles = [None] * sh.gc
        dist.all_gather_object(handles, (_export_tensor(local), _export_tensor(flags), self.epoch),
                               group=sh.half_group)
        me = sh.rank_in_half
        self._keep = (local, flags)                # the peers' mappings alias these allocations
        bufs, flag_views, err = [], [], None
        try:
            for r, (hb, hf, ep) in enumerate(handles):
                if ep != self.epoch:
                    raise RuntimeError("ranks diverged: peer exchange epochs differ (not every rank made the same "
                                       "calls)")
                b, f = (local, flags) if r == me else (_import_tensor(hb, self.device), _import_tensor(hf, self.device))
                bufs.append(b)
                flag_views.append(f.view(torch.int32).view(3, native.CSA_MAX_PEERS))
        except Exception as e:   # noqa: BLE001 - reported to every rank below, then raised everywhere
            err = f"rank {sh.rank}: {type(e).__name__}: {e}"
        # every rank learns whether every rank could map every buffer: either all go on or all raise (a rank that
        # raised alone would leave the others waiting in the next collective)
        errs = [None] * sh.gc
        dist.all_gather_object(errs, err, group=sh.half_group)
        failed = [e for e in errs if e]
        if failed:
            raise PeerExchangeUnavailable("peer-memory exchange could not be set up (" + "; ".join(failed) + "); "
                                          "use FrameSharding(exchange='nccl')")
        self.bufs, self.flags = bufs, flag_views
        self.cap_bytes = cap
        self._views = {}
        self._ready = [f[0] for f in self.flags]
        self._done = [f[1] for f in self.flags]
        self._counter = self.flags[me][2]
        self._counter2 = self.flags[me][2][1:]
        self.epoch = 0    # fresh flags: epochs restart
        self.allocations += 1
        dist.barrier(group=sh.half_group)          # every rank has mapped every buffer before the first store

    def end_step(self) -> None:
        """The step's exchange calls are over: fold their count into the device-resident base (kept even so that the
        slot of a call, epoch % SLOTS, is the same in every step)."""
        if self.epoch == 0 or self.epoch_base is None:
            return
        if self.epoch % self.SLOTS:
            # an odd step: pad it with an epoch that moves no data but IS released (a later scatter into that slot
            # waits for done >= its epoch - SLOTS, which may be this one)
            self.epoch += 1
            native.peer_signal(self._done, self.sh.rank_in_half, self.epoch, self.epoch_base,
                               epoch_base=self.epoch_base)
        native.epoch_advance(self.epoch_base, self.epoch)
        self.epoch = 0

    def views(self, slot: int, rows: int, cols: int, dtype):
        """Per rank the (K[S], V[S]) views [rows, cols] of ``slot`` (cached: the same few layer shapes recur)."""
        key = (slot, rows, cols, dtype)
        hit = self._views.get(key)
        if hit is not None:
            return hit
        nbytes = rows * cols * torch.empty((), dtype=dtype).element_size()
        if nbytes > self.cap_bytes:
            raise ValueError("PeerExchange.ensure() was not called for this layer size")
        ks, vs = [], []
        for b in self.bufs:
            k0 = (slot * 2) * self.cap_bytes
            v0 = (slot * 2 + 1) * self.cap_bytes
            ks.append(b[k0:k0 + nbytes].view(dtype).view(rows, cols))
            vs.append(b[v0:v0 + nbytes].view(dtype).view(rows, cols))
        self._views[key] = (ks, vs)
        return ks, vs

    def ready(self):
        return self._ready

    def done(self):
        return self._done

    def counter(self):
        return self._counter

    def counter2(self):
        """Second local counter (the attention launch's own release)."""
        return self._counter2


class ShardPlan:
    """Per (mask, sharding): the device-resident sampled list and its per-frame runs; on demand (NCCL path and
    reporting only — it costs a host read-back) who sends how many sampled rows and where they land."""

    def __init__(self, cm, sh: "FrameSharding", device):
        self.cm, self.sh, self.device = cm, sh, device
        self.s_idx, self.s_count, self.ranges = cm.sample_list(device)
        self._host = None

    def host(self) -> "ShardPlan":
        if self._host is None:
            sh, device = self.sh, self.device
            F, N = self.cm.id_length, self.cm.n_tokens
            rh = self.ranges.cpu().tolist()                # host sync
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
            self.local_idx = (self.s_idx[self.lo[me]:self.hi[me]] - sh.f0 * N).contiguous()
            # compact position i of S -> row of the all-gathered slab matrix [(rank, K|V, pad rows), C]
            parts = [torch.arange(c, dtype=torch.int32, device=device) + 2 * r * self.pad
                     for r, c in enumerate(self.counts)]
            self.slab_map = torch.cat(parts) if parts else torch.zeros((0,), dtype=torch.int32, device=device)
            self._host = True
        return self


class FrameSharding:
    """Static assignment of (CFG half, frame) units to the ranks of ``group`` (default: the world)."""

    def __init__(self, id_length: int, group=None, device: Optional[torch.device] = None, exchange: str = "p2p"):
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' (fused gather + peer-memory scatter) or 'nccl' (all-gather)")
        self.exchange = exchange
        self.mask_sync = "broadcast"   # or "seeded": see sync_sample()
        # SMs the attention launch leaves to a CONCURRENT peer scatter on a side stream (0 = the scatter runs before
        # the attention launch on the same stream).  CSA_OVERLAP_SMS overrides (tuning).
        self.overlap_sms = int(os.environ.get("CSA_OVERLAP_SMS", "0"))
        # p2p exchange fused into the K|V projection's epilogue and the release into the attention launch (three
        # launches per layer instead of five); CSA_FUSED_EXCHANGE=0: separate csa_peer_scatter_kv / csa_peer_signal
        self.fused_exchange = os.environ.get("CSA_FUSED_EXCHANGE", "1") != "0"
        self._side_streams = {}
        self.peers: Optional[PeerExchange] = None
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
        self._bytes_exchanged = 0                 # received bytes, for reporting
        self._p2p_sent = {}                       # id(plan) -> [plan, bytes per sampled row sent so far] (lazy)

    @property
    def bytes_exchanged(self) -> int:
        """Bytes this rank sent to its peers so far (reporting; reads the sampled counts back, so not for hot loops)."""
        total = self._bytes_exchanged
        for plan, per_row in self._p2p_sent.values():
            total += plan.host().count_me * per_row
        return total

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

    def sync_sample(self, sample: torch.Tensor) -> torch.Tensor:
        """``post_sample`` hook of ``masks.cal_attn_mask_xl``: returns the device tensor that holds rank 0's draw of
        a sample vector (in place when it already lives on this rank's device).  ``mask_sync="seeded"``: the ranks
        seed the generator alike (as ``setup_seed`` does), nothing is sent — the mode a captured step uses."""
        if self.device is not None and sample.device != torch.device(self.device):
            sample = sample.to(self.device)
        if self.mask_sync == "broadcast":
            dist.broadcast(sample.view(torch.uint8),
                           src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        return sample

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

    def can_fuse_exchange(self, cm) -> bool:
        """True when the K|V projection can deliver the sampled rows to the peers itself (``begin_exchange``)."""
        return (self.gc > 1 and self.exchange == "p2p" and self.fused_exchange and self.overlap_sms == 0
                and cm.shared_sample)

    def begin_exchange(self, cm, N: int, C: int, dtype, device, Fl: int) -> dict:
        """Open the exchange of one layer call BEFORE its K|V projection: returns what ``native.gemm`` needs to store
        this rank's sampled rows into every peer's K[S] / V[S] buffer from its epilogue (``scatter`` positions of the
        local rows, ``exchange`` dict) and what ``attn_write(..., exchanged=ctx)`` needs afterwards."""
        if Fl != self.id_length:
            raise ValueError(f"sharding was built for id_length {self.id_length}, processor has {Fl}")
        fr = self.frames_local
        me = self.rank_in_half
        pl = self.plan(cm, device)
        if self.peers is None:
            self.peers = PeerExchange(self, device)
        ex = self.peers
        rows = Fl * N + native.CSA_TILE
        ex.ensure(rows * C * torch.empty((), dtype=dtype).element_size())
        ex.epoch += 1
        epoch = ex.epoch
        ks, vs = ex.views(epoch % ex.SLOTS, rows, C, dtype)
        pos = cm.sample_positions(device)[self.f0 * N:(self.f0 + fr) * N]    # local row -> position in S (or -1)
        sent = self._p2p_sent.setdefault(id(pl.cm), [pl, 0])
        sent[0] = pl
        sent[1] += (self.gc - 1) * 2 * C * torch.empty((), dtype=dtype).element_size()
        return {"epoch": epoch, "ks": ks, "vs": vs, "rows": rows, "pos": pos, "plan": pl,
                "exchange": {"k_dst": ks, "v_dst": vs, "ready": ex.ready(), "self": me, "epoch": epoch,
                             "done": ex.done()[me], "done_epoch": epoch - ex.SLOTS, "counter": ex.counter(),
                             "epoch_base": ex.epoch_base}}

    def attn_write(self, q, k, v, o, N, heads, cm, Fl, exchanged: Optional[dict] = None):
        """Write-mode consistent attention of the local frames (``__call1__``, Comic_Generation.py:129-196 with
        mask[:F*N,:F*N]) with the sampled rows of the other ranks' frames fetched over NVLink.  ``exchanged``: the
        context of ``begin_exchange`` when the K|V projection has already delivered this rank's rows."""
        if Fl != self.id_length:
            raise ValueError(f"sharding was built for id_length {self.id_length}, processor has {Fl}")
        fr = self.frames_local
        if q.shape[0] != fr * N:
            raise ValueError(f"rank {self.rank} expects {fr} local frames ({fr * N} rows), got {q.shape[0]} rows")
        if not cm.shared_sample:
            raise ValueError("sharded consistent attention needs masks whose rows share one sample vector")
        C = q.shape[1]
        pl = self.plan(cm, q.device)
        if exchanged is not None and self.gc == 1:
            # one CFG half per GPU and the K|V projection has already filled K[S] / V[S] from its epilogue
            k_s, v_s, cap = exchanged
            native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=fr, n_q=N,
                            k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=pl.ranges, range_base=self.f0, range_step=1,
                            k_b=k, v_b=v, b_group_rows=fr * N, cb=(0, N, N))
            return o
        if exchanged is not None:
            ex = self.peers
            me = self.rank_in_half
            epoch = exchanged["epoch"]
            native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=fr, n_q=N,
                            k_a=exchanged["ks"][me], v_a=exchanged["vs"][me], a_group_rows=exchanged["rows"],
                            ranges=pl.ranges, range_base=self.f0, range_step=1, k_b=k, v_b=v, b_group_rows=fr * N,
                            cb=(0, N, N), b_first=True, ready=ex.ready()[me], ready_epoch=epoch, ready_peers=self.gc,
                            ready_frames_per_peer=fr, epoch_base=ex.epoch_base,
                            done=ex.done(), done_counter=ex.counter2(), peer_self=me)
            return o
        if self.gc > 1 and self.exchange == "p2p":
            return self._attn_write_p2p(q, k, v, o, N, heads, pl, Fl)
        if self.gc == 1:
            # one CFG half per GPU: everything is local, same as the single-GPU path with one group
            k_s, v_s, cap = native.gather_kv(k, v, Fl * N, 1, pl.s_idx, pl.s_count, Fl * N)
        else:
            pl.host()                   # slab sizes of the all-gather: one read-back per mask
            native.flush_batch()        # torch collectives follow: nothing deferred may be overtaken
            send = torch.empty((2, pl.pad, C), dtype=k.dtype, device=k.device)
            if pl.count_me > 0:
                native.gather_rows(k, pl.local_idx, pl.count_me, out=send[0])
                native.gather_rows(v, pl.local_idx, pl.count_me, out=send[1])
            recv = torch.empty((self.gc, 2, pl.pad, C), dtype=k.dtype, device=k.device)
            dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.half_group)
            self._bytes_exchanged += (self.gc - 1) * send.numel() * send.element_size()
            flat = recv.view(-1, C)
            # slab rows -> the compact K[S], V[S] buffers of the single-GPU path (zero tail included)
            k_s, v_s, cap = native.gather_kv(flat[:-pl.pad], flat[pl.pad:], flat.shape[0] - pl.pad, 1, pl.slab_map,
                                             pl.s_count, Fl * N)
        native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=fr, n_q=N,
                        k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=pl.ranges, range_base=self.f0, range_step=1,
                        k_b=k, v_b=v, b_group_rows=fr * N, cb=(0, N, N))
        return o

    def prepare_peers(self, layers, element_size: int = 2) -> None:
        """Collective over the CFG half: allocate and map the exchange buffers up front for the given layer shapes
        (iterable of ``(tokens per frame, channels)``).  Raises ``PeerExchangeUnavailable`` on every rank of the half
        if any of them cannot map its peers."""
        if self.gc > 1 and self.exchange == "p2p":
            if self.peers is None:
                self.peers = PeerExchange(self, self.device)
            self.peers.ensure(max((self.id_length * n + native.CSA_TILE) * c * element_size for (n, c) in layers))

    def _attn_write_p2p(self, q, k, v, o, N, heads, pl: ShardPlan, Fl):
        """Fused exchange: this rank's sampled rows go straight into every peer's S-ordered K[S], V[S] buffer
        (one kernel), the attention launch starts on the frames' own blocks and picks up each peer's rows when
        their flag is up, and a last tiny kernel tells the peers that the buffers of this epoch were read."""
        C = q.shape[1]
        fr = self.frames_local
        me = self.rank_in_half
        if self.peers is None:
            self.peers = PeerExchange(self, q.device)
        ex = self.peers
        rows = Fl * N + native.CSA_TILE            # S has at most F*N rows; a ragged last tile may overhang
        ex.ensure(rows * C * k.element_size())
        ex.epoch += 1
        epoch = ex.epoch
        ks, vs = ex.views(epoch % ex.SLOTS, rows, C, k.dtype)
        # Geometry stays on the device (no read-back of the sampled counts): the scatter kernel finds this rank's run of
        # S in `ranges`, the attention kernel the peers' bounds.  Peers last read this slot in epoch - SLOTS.
        #
        # The scatter runs on a SIDE stream, concurrently with the attention launch: the attention kernel does not
        # depend on the scatter's completion but on the arrival flags (its own rows' flag included), starts on each
        # frame's own — local — block, and leaves `overlap_sms` SMs free so that the scatter's CTAs always have
        # somewhere to run (a persistent attention launch that filled every SM while waiting for flags would
        # otherwise starve the very kernel that raises them).  Measured: the scatter is 21-35 us (F=4) / 58-104 us
        # (F=16) per layer over real links (profiles/r02v_peer_scatter_nvlink.md), ~10 % of a step when exposed.
        side = None
        if self.overlap_sms > 0 and q.is_cuda:
            native.flush_batch()                   # the projections are on the main stream from here
            main = torch.cuda.current_stream(q.device)
            side = self._side_stream(q.device)
            side.wait_stream(main)
        ctx = torch.cuda.stream(side) if side is not None else _NULL
        with ctx:
            native.peer_scatter_kv(k, v, pl.s_idx, fr * N, 0, ks, vs, ex.ready(), me, epoch,
                                   ex.done()[me], epoch - ex.SLOTS, ex.counter(),
                                   ranges=pl.ranges, frames_per_peer=fr, idx_adjust=-self.f0 * N,
                                   epoch_base=ex.epoch_base)
        sent = self._p2p_sent.setdefault(id(pl.cm), [pl, 0])   # per mask object (masks are re-sampled in place)
        sent[0] = pl
        sent[1] += (self.gc - 1) * 2 * C * k.element_size()
        native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=fr, n_q=N,
                        k_a=ks[me], v_a=vs[me], a_group_rows=rows, ranges=pl.ranges, range_base=self.f0,
                        range_step=1, k_b=k, v_b=v, b_group_rows=fr * N, cb=(0, N, N),
                        b_first=True, ready=ex.ready()[me], ready_epoch=epoch, ready_peers=self.gc,
                        ready_frames_per_peer=fr, epoch_base=ex.epoch_base,
                        max_ctas=(self._sm_count(q.device) - self.overlap_sms) if side is not None else 0)
        if side is not None:
            native.flush_batch()
            torch.cuda.current_stream(q.device).wait_stream(side)   # join: k / v may be reused after this call
        native.peer_signal(ex.done(), me, epoch, q, epoch_base=ex.epoch_base)
        return o

    def _side_stream(self, device):
        st = self._side_streams.get(device.index)
        if st is None:
            st = self._side_streams[device.index] = torch.cuda.Stream(device)
        return st

    @staticmethod
    def _sm_count(device) -> int:
        return torch.cuda.get_device_properties(device).multi_processor_count

    def end_step(self) -> None:
        """Called by the processor at the step roll-over (every rank makes the same calls)."""
        if self.peers is not None:
            self.peers.end_step()

    # ---------------------------------------------------------------------------------------------- the bank
    def gather_bank(self, k: torch.Tensor, v: torch.Tensor):
        """All-gather the K/V rows a sharded write pass left on this rank (its own frames of its CFG half) into the
        full id_bank layout ``[uncond frames 0..F-1, cond frames 0..F-1]`` — rank order is exactly that order — so
        that read passes (Comic_Generation.py:441-448: one generated frame per call, independent given the bank) can
        run on any rank, different frames on different ranks.  One collective per (layer, step), done lazily by the
        first read of that entry; the result replaces the shard in the bank."""
        native.flush_batch()
        rows, C = k.shape
        ks = k.contiguous()
        vs = v.contiguous()
        kf = torch.empty((self.world * rows, C), dtype=k.dtype, device=k.device)
        vf = torch.empty_like(kf)
        dist.all_gather_into_tensor(kf.view(-1), ks.view(-1), group=self.group)
        dist.all_gather_into_tensor(vf.view(-1), vs.view(-1), group=self.group)
        self._bytes_exchanged += 2 * (self.world - 1) * ks.numel() * ks.element_size()
        return kf, vf
