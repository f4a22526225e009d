"""CPU, gloo: the (CFG half, frame) sharding of spider_b200/dist.py with 2 and 4 ranks.  The native entry points are
replaced by the torch emulations of tests/abi_emulation.py (test doubles following the C-ABI semantics), so what is
checked is the host logic: which rank owns which frames, the run bookkeeping, the all-gather layout, the compaction
of the slabs into the K[S], V[S] buffers, mask broadcast and lock-step — every rank's outputs must equal the rows a
single unsharded processor produces for the same frames, and both must equal the reference oracle."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import spider_b200
from spider_b200 import native
from spider_b200.install import make_processor_class
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention

import abi_emulation

H = W = 128            # N = 16 at /32, 64 at /16
FL, C, HEADS = 4, 128, 2
STEPS = 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_inputs():
    g = torch.Generator().manual_seed(7)
    attn = FakeAttention(C, HEADS)
    with torch.no_grad():
        for p in attn.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    xs = [[torch.randn((2 * FL, n, C), generator=g) for n in (16, 64)] for _ in range(STEPS)]
    return attn, xs


def _run_story(host, procs, attn, xs, slicer):
    """STEPS denoise steps x 2 layers (one per resolution), consistent branch forced by the seed; returns outputs."""
    outs = []
    random.seed(11)
    torch.manual_seed(5)
    host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(FL + 1, FL, 0.5, 0.5, H, W, "cpu", torch.float32)
    if procs[0].dist is not None:
        procs[0].dist.sync_masks(host.mask1024, host.mask4096)
    host.write, host.cur_step, host.attn_count = True, 25, 0
    orig = random.random
    random.random = lambda: 0.5
    try:
        with torch.no_grad():
            for s in range(STEPS):
                for li, p in enumerate(procs):
                    outs.append(p(attn, slicer(xs[s][li])))
    finally:
        random.random = orig
    return outs


def _worker(rank, world, port, q, exchange="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    if exchange == "p2p":
        mp.set_sharing_strategy("file_system")   # peer buffers of the exchange are shared-memory CPU tensors here
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spider_b200.dist import FrameSharding

        host = spider_b200.StoryGlobals()
        host.height, host.width, host.total_count = H, W, 2
        cls = make_processor_class(host)
        abi_emulation.install(None, native, cls)
        sh = FrameSharding(FL, None, torch.device("cpu"), exchange=exchange)
        procs = [cls(id_length=FL, device="cpu", dtype=torch.float32) for _ in range(2)]
        for p in procs:
            p.dist = sh
        attn, xs = _make_inputs()
        if rank != 0:
            torch.manual_seed(1000 + rank)   # a rank whose torch generator drifted: masks must still agree

        def slicer(x):   # the latents this rank's UNet replica would carry
            half = x[sh.cfg * FL:(sh.cfg + 1) * FL]
            return half[sh.f0:sh.f0 + sh.frames_local].contiguous()

        real_seed = torch.manual_seed

        def seed_only_rank0(s):    # _run_story re-seeds torch: keep the drift on the other ranks
            return real_seed(s if rank == 0 else s + 1000 + rank)

        torch.manual_seed = seed_only_rank0
        try:
            outs = _run_story(host, procs, attn, xs, slicer)
        finally:
            torch.manual_seed = real_seed
        sh.check_lockstep(0.5)
        q.put((rank, sh.cfg, sh.f0, sh.frames_local, [o.numpy().copy() for o in outs], sh.bytes_exchanged,
               sorted(procs[0].id_bank.keys()), list(abi_emulation.PEER_LOG),
               (sh.peers.allocations, sh.peers.epoch) if sh.peers is not None else None))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange", [(2, "nccl"), (4, "nccl"), (4, "p2p"), (8, "p2p")])
def test_sharded_write_pass_matches_single_process_and_oracle(world, exchange, monkeypatch):
    # single process, unsharded, same emulated ABI
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 2
    cls = make_processor_class(host)
    abi_emulation.install(monkeypatch, native, None)
    monkeypatch.setattr(cls, "_check_input", staticmethod(lambda x: None))
    procs = [cls(id_length=FL, device="cpu", dtype=torch.float32) for _ in range(2)]
    attn, xs = _make_inputs()
    single = _run_story(host, procs, attn, xs, lambda x: x)

    # the reference oracle on the same inputs, masks and gate
    st = rp.StoryState(write=True, cur_step=25, total_count=2, height=H, width=W)
    random.seed(11)
    torch.manual_seed(5)
    st.mask1024, st.mask4096 = rp.cal_attn_mask_xl(FL + 1, FL, 0.5, 0.5, H, W)
    orcs = [rp.ConsistentAttnOracle(st, id_length=FL) for _ in range(2)]
    orig = random.random
    random.random = lambda: 0.5
    want = []
    try:
        with torch.no_grad():
            for s in range(STEPS):
                for li, o in enumerate(orcs):
                    want.append(o(attn, xs[s][li]))
    finally:
        random.random = orig
    for a, b in zip(single, want):
        assert (a - b).abs().max().item() < 2e-5

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, exchange)) for r in range(world)]
    for p in ps:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    gc = world // 2
    seen = set()
    for rank, cfg, f0, fr, outs, nbytes, bank_keys, peer_log, peer_state in results:
        assert cfg == rank // gc and fr == FL // gc and f0 == (rank % gc) * fr
        seen.add((cfg, f0))
        for got, ref in zip(outs, single):
            ref_rows = ref[cfg * FL + f0:cfg * FL + f0 + fr]
            got = torch.from_numpy(got)
            assert got.shape == ref_rows.shape
            assert (got - ref_rows).abs().max().item() < 2e-5
        assert (nbytes > 0) == (gc > 1)          # G == 2 exchanges nothing
        assert bank_keys == [25, 26, 27]
        if exchange == "p2p" and gc > 1:
            # one fused scatter + one release signal per consistent layer call.  Epochs live in device memory: a call
            # passes its number within the step, the kernels add the base, end_step() advances the base by the (even)
            # number of calls — an odd step is padded with an epoch that is only released.  The smaller layer comes
            # first, so the buffers are re-allocated by the second call (epochs restart there: step 0 is 1 | 1 + pad).
            n_calls = STEPS * 2
            real = [e for e in peer_log if e[0] in ("scatter", "signal")]
            assert [e[0] for e in real] == ["scatter", "signal", "scatter", "signal", "signal"] + \
                ["scatter", "signal"] * (n_calls - 2)
            assert [e for e in peer_log if e[0] == "advance"] == [("advance", 2)] * STEPS
            assert peer_state == (2, 0)
            epochs = [e[1] for e in peer_log if e[0] == "scatter"]
            assert epochs == [1, 1] + list(range(3, n_calls + 1))
            assert [e[1] for e in real if e[0] == "signal"] == [1, 1, 2] + list(range(3, n_calls + 1))
            # a slot is rewritten only after every peer has released it: done_epoch = epoch - 2 (no wait if <= 0)
            assert all(e[4] == e[1] - 2 for e in peer_log if e[0] == "scatter")
        else:
            assert peer_log == [] and peer_state is None
    assert len(seen) == world


def _worker_fused(rank, world, port, q):
    """The fused exchange of the GPU path, driven at the dist layer (on a GPU the processor does exactly this around
    its K|V projection): begin_exchange -> gemm(exchange=...) -> attn_write(exchanged=...), three steps of two layer
    shapes, against the unsharded computation of the same inputs."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    mp.set_sharing_strategy("file_system")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spider_b200.dist import FrameSharding

        abi_emulation.install(None, native, None)
        sh = FrameSharding(FL, None, torch.device("cpu"), exchange="p2p")
        assert sh.fused_exchange
        random.seed(11)
        torch.manual_seed(5)
        m1, m2 = spider_b200.cal_attn_mask_xl(FL + 1, FL, 0.5, 0.5, H, W, "cpu", torch.float32)
        sh.sync_masks(m1, m2)
        g = torch.Generator().manual_seed(123)
        w_qkv = torch.randn((3 * C, C), generator=g) * 0.05
        fr, f0 = sh.frames_local, sh.f0
        worst = 0.0
        for step in range(STEPS):
            for cm, n in ((m1, 16), (m2, 64)):
                x_full = torch.randn((2 * FL * n, C), generator=g)           # both halves, all frames
                # ---- unsharded: project, gather, attend (the single-GPU fused path through the same emulation)
                qf, kvf = torch.empty((2 * FL * n, C)), torch.empty((2 * FL * n, 2 * C))
                cap = FL * n + native.CSA_TILE
                k_s, v_s = torch.zeros((2 * cap, C)), torch.zeros((2 * cap, C))
                pos = cm.sample_positions("cpu")
                native.gemm(x_full, w_qkv, out=qf, out2=kvf, scatter=(pos, k_s, v_s, FL * n, cap, 2 * C, C))
                s_idx, s_count, ranges = cm.sample_list("cpu")
                want = torch.empty_like(qf)
                native.attn_fwd(qf, want, heads=HEADS, n_groups=2, n_frames=FL, n_q=n, k_a=k_s, v_a=v_s,
                                a_group_rows=cap, ranges=ranges, range_base=0, range_step=1, k_b=kvf[:, :C],
                                v_b=kvf[:, C:], b_group_rows=FL * n, cb=(0, n, n))
                # ---- sharded: this rank's frames of its half
                r0 = (sh.cfg * FL + f0) * n
                x = x_full[r0:r0 + fr * n].contiguous()
                q_l, kv_l = torch.empty((fr * n, C)), torch.empty((fr * n, 2 * C))
                assert sh.can_fuse_exchange(cm)
                ctx = sh.begin_exchange(cm, n, C, torch.float32, torch.device("cpu"), FL)
                native.gemm(x, w_qkv, out=q_l, out2=kv_l, scatter=(ctx["pos"], None, None, fr * n, 0, 2 * C, C),
                            exchange=ctx["exchange"])
                o = torch.empty_like(q_l)
                sh.attn_write(q_l, kv_l[:, :C], kv_l[:, C:], o, n, HEADS, cm, FL, exchanged=ctx)
                worst = max(worst, (o - want[r0:r0 + fr * n]).abs().max().item())
            sh.end_step()
            m1.resample_(0.5, torch.float32, post_sample=sh.sync_sample)
            m2.resample_(0.5, torch.float32, post_sample=sh.sync_sample)
        q.put((rank, worst, list(abi_emulation.PEER_LOG), (sh.peers.allocations, sh.peers.epoch)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [4, 8])
def test_fused_exchange_protocol(world):
    """csa_gemm(exchange=...) + csa_attn_fwd(done_dst=...): the K|V projection delivers the sampled rows to every peer
    and raises the arrival flags, the attention launch releases the buffers itself — no csa_peer_scatter_kv and no
    csa_peer_signal launch (except the pad of an odd step), same epochs and slot discipline as the unfused path."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_fused, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_calls = STEPS * 2
    for rank, worst, peer_log, peer_state in results:
        assert worst < 2e-5, (rank, worst)
        kinds = [e[0] for e in peer_log if e[0] != "advance"]
        # the smaller layer comes first: the buffers are re-allocated by the second call, epochs restart there, and the
        # (then odd) first step is padded with an epoch that is only released (PeerExchange.end_step)
        assert kinds == ["exchange", "release", "exchange", "release", "signal"] + ["exchange", "release"] * (n_calls - 2)
        assert "scatter" not in kinds
        assert [e for e in peer_log if e[0] == "advance"] == [("advance", 2)] * STEPS
        ex = [e for e in peer_log if e[0] == "exchange"]
        assert [e[1] for e in ex] == [1, 1] + list(range(3, n_calls + 1))
        assert all(e[3] == e[1] - 2 for e in ex)                       # slot reuse waits for epoch - 2
        assert [e[1] for e in peer_log if e[0] == "release"] == [e[1] for e in ex]
        assert peer_state == (2, 0)


def _read_inputs(world):
    g = torch.Generator().manual_seed(77)
    return [[[torch.randn((2, n, C), generator=g) for n in (16, 64)] for _ in range(STEPS)] for _ in range(world)]


def _run_reads(host, procs, attn, xr):
    """one generated frame through the read pass of the STEPS written steps (Comic_Generation.py:441-448)"""
    outs = []
    host.write, host.cur_step, host.attn_count = False, 25, 0
    orig = random.random
    random.random = lambda: 0.5
    try:
        with torch.no_grad():
            for s in range(STEPS):
                for li, p in enumerate(procs):
                    outs.append(p(attn, xr[s][li]))
    finally:
        random.random = orig
    return outs


def _worker_story(rank, world, port, q, exchange):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    if exchange == "p2p":
        mp.set_sharing_strategy("file_system")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spider_b200.dist import FrameSharding

        host = spider_b200.StoryGlobals()
        host.height, host.width, host.total_count = H, W, 2
        cls = make_processor_class(host)
        abi_emulation.install(None, native, cls)
        sh = FrameSharding(FL, None, torch.device("cpu"), exchange=exchange)
        procs = [cls(id_length=FL, device="cpu", dtype=torch.float32) for _ in range(2)]
        for p in procs:
            p.dist = sh
        attn, xs = _make_inputs()

        def slicer(x):
            half = x[sh.cfg * FL:(sh.cfg + 1) * FL]
            return half[sh.f0:sh.f0 + sh.frames_local].contiguous()

        _run_story(host, procs, attn, xs, slicer)                      # sharded write pass: sharded bank
        shard_rows = [int(p.id_bank[25].k.shape[0]) for p in procs]
        outs = _run_reads(host, procs, attn, _read_inputs(world)[rank])  # every rank generates ITS OWN frame
        whole_rows = [int(p.id_bank[25].k.shape[0]) for p in procs]
        q.put((rank, [o.numpy().copy() for o in outs], shard_rows, whole_rows))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange", [(2, "nccl"), (4, "p2p")])
def test_sharded_story_finishes_write_then_frame_parallel_reads(world, exchange, monkeypatch):
    """An N-GPU story runs to the end: sharded write pass, the sharded id_bank is all-gathered by the first read of
    each (layer, step) entry, and every rank then generates a different frame — each equal to what the single-process
    processor produces for that frame from the unsharded bank."""
    import copy

    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 2
    cls = make_processor_class(host)
    abi_emulation.install(monkeypatch, native, None)
    monkeypatch.setattr(cls, "_check_input", staticmethod(lambda x: None))
    procs = [cls(id_length=FL, device="cpu", dtype=torch.float32) for _ in range(2)]
    attn, xs = _make_inputs()
    _run_story(host, procs, attn, xs, lambda x: x)
    rng_after_write = torch.get_rng_state()
    masks_after_write = copy.deepcopy((host.mask1024, host.mask4096))
    want = []
    for r in range(world):
        torch.set_rng_state(rng_after_write)
        host.mask1024, host.mask4096 = copy.deepcopy(masks_after_write)
        want.append(_run_reads(host, procs, attn, _read_inputs(world)[r]))

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_story, args=(r, world, port, q, exchange)) for r in range(world)]
    for p in ps:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, outs, shard_rows, whole_rows in results:
        fr = FL // (world // 2)
        assert shard_rows == [fr * 16, fr * 64] and whole_rows == [2 * FL * 16, 2 * FL * 64]
        assert len(outs) == len(want[rank])
        for got, ref in zip(outs, want[rank]):
            assert (torch.from_numpy(got) - ref).abs().max().item() < 2e-5


def test_sharding_geometry_errors():
    class _G:   # get_world_size / get_rank are looked up on torch.distributed: exercise the checks through a stub
        pass
    import spider_b200.dist as sd

    def fake(world, rank):
        class D:
            @staticmethod
            def is_initialized():
                return True

            @staticmethod
            def get_world_size(group=None):
                return world

            @staticmethod
            def get_rank(group=None):
                return rank

            @staticmethod
            def new_group(ranks=None):
                return tuple(ranks)

            @staticmethod
            def get_process_group_ranks(group):
                return list(range(world))
        return D

    real = sd.dist
    try:
        sd.dist = fake(3, 0)
        with pytest.raises(ValueError, match="even number"):
            sd.FrameSharding(4)
        sd.dist = fake(8, 5)
        with pytest.raises(ValueError, match="not divisible"):
            sd.FrameSharding(6)
        sh = sd.FrameSharding(16)
        assert (sh.gc, sh.cfg, sh.rank_in_half, sh.frames_local, sh.f0) == (4, 1, 1, 4, 4)
        assert sh.half_group == (4, 5, 6, 7)
        sd.dist = fake(2, 1)
        sh = sd.FrameSharding(4)
        assert (sh.gc, sh.cfg, sh.frames_local, sh.f0, sh.half_group) == (1, 1, 4, 0, None)
    finally:
        sd.dist = real


def _worker_unavailable(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    mp.set_sharing_strategy("file_system")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spider_b200.dist as sd
        if rank == 1:     # one rank of the unconditional half cannot map its peer
            def broken(handle, device):
                raise OSError("peer mapping refused")
            sd._import_tensor = broken
        sh = sd.FrameSharding(FL, None, torch.device("cpu"), exchange="p2p")
        try:
            sh.prepare_peers([(16, C), (64, C)], element_size=4)
            q.put((rank, "ok", sh.peers.allocations))
        except sd.PeerExchangeUnavailable as e:
            q.put((rank, "unavailable", str(e)))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_peer_exchange_failure_is_seen_by_the_whole_half():
    """A rank that cannot map a peer's buffers must not leave its peers waiting: every rank of that CFG half raises
    PeerExchangeUnavailable (and the other half, which is independent, goes on)."""
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_unavailable, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    results = dict((r[0], r[1:]) for r in [q.get(timeout=240) for _ in range(world)])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][0] == results[1][0] == "unavailable" and "rank 1" in results[0][1]
    assert results[2] == ("ok", 1) and results[3] == ("ok", 1)
