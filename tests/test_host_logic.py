"""CPU: the host-side state machine of the B200 processor (branching, RNG consumption, counters, bank, mask
handling, install/uninstall) with the native calls replaced by recorders.  The recorders are test doubles that
capture the launch geometry — they compute nothing; numerical parity is covered by the gpu tests."""
import copy
import random
import types

import numpy as np
import pytest
import torch

import spider_b200
from spider_b200 import masks as csa_masks
from spider_b200 import native
from spider_b200.bank import BankEntry, IdBank
from spider_b200.install import install, make_processor_class, set_attention_processor, uninstall
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention, FakeUNet, sdxl_layout

from helpers import load_npz


@pytest.fixture
def recorder(monkeypatch):
    calls = []

    def fake_attn_fwd(q, o, **kw):
        o.zero_()
        calls.append(("attn", {k: v for k, v in kw.items() if not isinstance(v, torch.Tensor)},
                      {k: tuple(v.shape) for k, v in kw.items() if isinstance(v, torch.Tensor)}, tuple(q.shape)))
        return o

    def fake_compact_rows(mask_rows, n_rows, n_cols, row_stride, block_n=0, limit_cols=0, idx=None, counts=None):
        calls.append(("compact", n_rows, n_cols, row_stride, block_n, limit_cols))
        if idx is not None and counts is not None:    # in-place refresh of a re-sampled mask
            return idx, counts
        return (torch.zeros((n_rows, native.idx_stride_for(n_cols)), dtype=torch.int32),
                torch.zeros((n_rows,), dtype=torch.int32))

    def fake_validate(mask, block_n):
        calls.append(("validate", tuple(mask.shape), block_n))
        return torch.zeros((1,), dtype=torch.int32)

    def fake_sample_ranges(s_idx, s_count, block_n, n_frames, out=None):
        calls.append(("ranges", block_n, n_frames))
        return torch.zeros((n_frames + 1, 4), dtype=torch.int32) if out is None else out

    def fake_gather_kv(k, v, group_rows, n_groups, s_idx, s_count, max_rows):
        calls.append(("gather_kv", tuple(k.shape), group_rows, n_groups, max_rows))
        cap = max_rows + native.CSA_TILE
        return (torch.zeros((n_groups * cap, k.shape[1]), dtype=k.dtype),
                torch.zeros((n_groups * cap, k.shape[1]), dtype=k.dtype), cap)

    monkeypatch.setattr(native, "sample_ranges", fake_sample_ranges)
    monkeypatch.setattr(native, "gather_kv", fake_gather_kv)
    monkeypatch.setattr(native, "attn_fwd", fake_attn_fwd)
    monkeypatch.setattr(native, "compact_rows", fake_compact_rows)
    monkeypatch.setattr(native, "validate_mask", fake_validate)
    monkeypatch.setattr(spider_b200.SpatialAttnProcessor2_0, "_check_input", staticmethod(lambda x: None))
    return calls


def _host(H=64, W=64, total=3):
    h = types.SimpleNamespace(write=False, cur_step=0, attn_count=0, total_count=total, sa32=0.5, sa64=0.5,
                              height=H, width=W, mask1024=None, mask4096=None)
    return h


def test_state_machine_matches_reference_trace(recorder):
    """Same scenario as the golden story: gate draws, branch per call, step roll-over, bank."""
    z = load_npz("story.npz")
    H, W, Fl, C, heads, steps = (int(x) for x in z["geom"])
    host = _host(H, W, 3)
    cls = make_processor_class(host)
    rp.setup_seed(2047)
    attns = [FakeAttention(C, heads) for _ in range(3)]
    procs = copy.deepcopy([cls(id_length=Fl, device="cpu", dtype=torch.float32) for _ in range(3)])
    host.mask1024, host.mask4096 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)   # dense, like the driver (:376)
    draws = []
    orig = random.random
    branches = []
    try:
        random.random = lambda: (draws.append(orig()), draws[-1])[1]
        with torch.no_grad():
            for phase, write in (("w", True), ("r", False)):
                host.write, host.cur_step = write, 0
                for s in range(steps):
                    for li, p in enumerate(procs):
                        x = torch.from_numpy(z[f"{phase}{s}_{li}_in"])
                        out = p(attns[li], x)
                        assert out.shape == x.shape and out.dtype == x.dtype
                        branches.append((p._last_branch, s))
    finally:
        random.random = orig
    assert np.array_equal(np.array(draws), z["draws"])
    want = [str(t).split(":") for t in z["trace"]]
    assert [("consistent" if b == "consistent" else "standard") for b, _ in branches] == [w[0] for w in want]
    assert [s for _, s in branches] == [int(w[1]) for w in want]
    assert host.cur_step == int(z["final_cur_step"]) and host.attn_count == 0
    for p, keys in zip(procs, z["bank_keys"]):
        assert sorted(p.id_bank.keys()) == list(keys)
    # masks are regenerated in compact form after every step, with the reference's RNG consumption
    assert isinstance(host.mask1024, csa_masks.CompactMask) and isinstance(host.mask4096, csa_masks.CompactMask)
    # the regenerated sample vectors equal the oracle's under the same seed
    rp.setup_seed(2047)
    [FakeAttention(C, heads) for _ in range(3)]
    rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)
    for _ in range(2 * steps - 1):
        r32, r16 = rp.sample_vectors(Fl + 1, 0.5, 0.5, H, W)
    r32, r16 = rp.sample_vectors(Fl + 1, 0.5, 0.5, H, W)
    assert torch.equal(host.mask1024._sample, r32) and torch.equal(host.mask4096._sample, r16)


def _last_attn(calls):
    """the most recent attention launch (a step roll-over re-samples the masks in place AFTER it: compact / ranges)"""
    return next(c for c in reversed(calls) if c[0] == "attn")


def test_launch_geometry_per_branch(recorder):
    Fl, N, C, heads = 4, 16, 128, 2
    host = _host(128, 128, 1)     # n32 = 16 -> mask1024
    cls = make_processor_class(host)
    attn = FakeAttention(C, heads)
    p = cls(id_length=Fl, device="cpu", dtype=torch.float32)
    host.mask1024, host.mask4096 = csa_masks.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, 128, 128, "cpu", torch.float32)
    random.seed(3)   # first draws: 0.2379 (standard at thr 0.3), 0.544 (consistent)
    with torch.no_grad():
        host.write, host.cur_step = True, 0
        p(attn, torch.randn(2 * Fl, N, C))
        kind, kw, shapes, qshape = _last_attn(recorder)
        assert kw["n_groups"] == 1 and kw["n_frames"] == 2 * Fl and kw["cb"] == (0, N, N) and "k_a" not in shapes
        host.cur_step = 5
        p(attn, torch.randn(2 * Fl, N, C))
        assert p._last_branch == "standard"
        p(attn, torch.randn(2 * Fl, N, C))
        assert p._last_branch == "consistent"
        kind, kw, shapes, qshape = _last_attn(recorder)
        # default: sampled rows gathered once per layer, frame f attends two runs of that buffer + its own block
        cap = Fl * N + native.CSA_TILE
        assert (kw["n_groups"], kw["n_frames"], kw["n_q"]) == (2, Fl, N) and "list_base" not in kw
        assert (kw["range_base"], kw["range_step"], kw["cb"]) == (0, 1, (0, N, N))
        assert kw["a_group_rows"] == cap and shapes["k_a"] == (2 * cap, C) and shapes["k_b"] == (2 * Fl * N, C)
        assert shapes["ranges"] == (Fl + 1, 4)
        assert ("compact", 1, Fl * N, 0, 0, 0) in recorder and ("ranges", N, Fl) in recorder
        assert ("gather_kv", (2 * Fl * N, C), Fl * N, 2, Fl * N) in recorder
        # generic alternative: per-frame index lists, TMA gather4 inside the attention kernel
        host.cur_step = 5
        random.seed(4)   # 0.236 -> standard
        cls.kv_gather = "inline"
        random.seed(0)   # 0.844 -> consistent
        p(attn, torch.randn(2 * Fl, N, C))
        cls.kv_gather = "pre"
        kind, kw, shapes, qshape = _last_attn(recorder)
        assert (kw["n_groups"], kw["n_frames"], kw["n_q"], kw["list_base"], kw["list_step"]) == (2, Fl, N, 0, 1)
        assert kw["a_group_rows"] == Fl * N and shapes["k_a"] == (2 * Fl * N, C)
        assert ("compact", Fl + 1, (Fl + 1) * N, 0, N, Fl * N) in recorder
        # read pass
        host.write, host.cur_step = False, 0
        p(attn, torch.randn(2, N, C))
        kind, kw, shapes, qshape = _last_attn(recorder)
        assert kw["ca"] == (0, 0, Fl * N) and kw["cb"] == (0, N, N) and kw["n_frames"] == 1   # frame step N (batched read); single frame here
        assert shapes["k_a"] == (2 * Fl * N, C) and shapes["k_b"] == (2 * N, C)
        host.cur_step = 6   # written above (total_count == 1: every call advances the step)
        random.seed(1)  # 0.134 -> standard, then 0.847 -> consistent
        p(attn, torch.randn(2, N, C))
        assert p._last_branch == "standard" and "k_a" not in _last_attn(recorder)[2]
        host.cur_step = 6
        p(attn, torch.randn(2, N, C))
        kind, kw, shapes, qshape = _last_attn(recorder)
        assert p._last_branch == "consistent"
        assert (kw["range_base"], kw["range_step"], kw["cb"]) == (Fl, 0, (0, N, N)) and "list_base" not in kw
        assert shapes["k_b"] == (2 * N, C) and kw["a_group_rows"] == Fl * N + native.CSA_TILE
        host.cur_step = 6
        cls.kv_gather = "inline"
        random.seed(0)
        p(attn, torch.randn(2, N, C))
        cls.kv_gather = "pre"
        kind, kw, shapes, qshape = _last_attn(recorder)
        assert (kw["list_base"], kw["list_step"], kw["g_adjust"]) == (Fl, 0, -N) and kw["cb"] == (0, N, N)
    with pytest.raises(KeyError):
        host.cur_step = 99
        p(attn, torch.randn(2, N, C))


def test_error_behaviour(recorder):
    Fl, N, C, heads = 4, 16, 128, 2
    host = _host(128, 128, 1)
    cls = make_processor_class(host)
    attn = FakeAttention(C, heads)
    p = cls(id_length=Fl, device="cpu", dtype=torch.float32)
    with pytest.raises(NotImplementedError):
        p(attn, torch.randn(8, N, C), attention_mask=torch.ones(1))
    with pytest.raises(NotImplementedError):
        p(attn, torch.randn(8, N, C), encoder_hidden_states=torch.randn(8, 77, C))
    host.write, host.cur_step = True, 6
    host.mask1024, host.mask4096 = csa_masks.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, 256, 256, "cpu", torch.float32)
    random.seed(0)  # 0.844 -> consistent
    with pytest.raises(ValueError, match="tokens per frame"):
        p(attn, torch.randn(8, N, C))      # masks were sampled for another latent size
    host.mask1024, host.mask4096 = csa_masks.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, 128, 128, "cpu", torch.float32)
    random.seed(0)
    with pytest.raises(ValueError, match="2\\*id_length"):
        p(attn, torch.randn(6, N, C))      # predict.py-style num_ids != id_length


def test_dense_mask_is_validated_and_cached(recorder):
    Fl, N, C, heads = 4, 16, 128, 2
    host = _host(128, 128, 2)
    cls = make_processor_class(host)
    attn = FakeAttention(C, heads)
    procs = [cls(id_length=Fl, device="cpu", dtype=torch.float32) for _ in range(2)]
    host.mask1024, host.mask4096 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, 128, 128)
    host.write, host.cur_step = True, 6
    random.seed(0)   # 0.844, 0.757 -> both consistent
    with torch.no_grad():
        for p in procs:
            p(attn, torch.randn(8, N, C))
    assert sum(1 for c in recorder if c[0] == "validate") == 1
    assert sum(1 for c in recorder if c[0] == "compact") == 1
    comp = [c for c in recorder if c[0] == "compact"][0]
    assert comp[1:] == (1, Fl * N, 0, 0, 0)          # the shared sample list S = read row restricted to F*N columns
    cm = host._csa_dense_cache[True][1]
    assert cm.shared_sample
    # a per-frame mask that is NOT (S u own block) keeps the generic per-frame lists (in-kernel gather4)
    odd = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, 128, 128)[0].clone()
    odd[:N, 2 * N:3 * N] = ~odd[:N, 2 * N:3 * N]
    host.mask1024 = odd
    host.attn_count, host.cur_step = 0, 6
    del recorder[:]
    random.seed(0)
    with torch.no_grad():
        procs[0](attn, torch.randn(8, N, C))
    assert not host._csa_dense_cache[True][1].shared_sample
    comp = [c for c in recorder if c[0] == "compact"][0]
    assert comp[1:] == (Fl + 1, (Fl + 1) * N, N * (Fl + 1) * N, 0, 0)   # rows 0, N, 2N.. of the dense mask
    assert not any(c[0] == "gather_kv" for c in recorder)


def test_bank_entry_semantics():
    bank = IdBank()
    hu, hc = torch.randn(4, 16, 128), torch.randn(4, 16, 128)
    bank[3] = [hu, hc]                       # reference-style assignment from outside
    e = bank[3]
    assert isinstance(e, BankEntry) and e[0] is hu and e[1] is hc and not e.has_kv()
    attn = FakeAttention(128, 2)
    with torch.no_grad():
        k, v = e.kv(attn)
    assert k.shape == (2 * 4 * 16, 128) and e.has_kv()
    with torch.no_grad():
        want = attn.to_k(torch.cat((hu, hc)).reshape(-1, 128))
    assert torch.allclose(k, want)
    with pytest.raises(KeyError):
        bank[4]
    assert bank.nbytes() == (hu.numel() + hc.numel() + k.numel() + v.numel()) * 4


def test_install_binds_host_module_and_unet_surface(recorder):
    host = types.ModuleType("fake_host")
    host.SpatialAttnProcessor2_0 = object
    host.cal_attn_mask_xl = rp.cal_attn_mask_xl
    cls = install(host)
    assert host.SpatialAttnProcessor2_0 is cls and cls._host is host
    assert host.cal_attn_mask_xl is csa_masks.cal_attn_mask_xl
    assert host.write is False and host.cur_step == 0
    unet = FakeUNet(sdxl_layout(), with_cross=True)
    unet.set_attn_processor(object())   # something installed everywhere
    n = set_attention_processor(unet, id_length=4, host=host)
    assert n == 36 and host.total_count == 36          # SDXL up-blocks: 30 + 6
    names = [k for k, v in unet.attn_processors.items() if isinstance(v, cls)]
    assert len(names) == 36 and all(k.startswith("up_blocks") and k.endswith("attn1.processor") for k in names)
    n = set_attention_processor(unet, id_length=4, host=host, all_self_attn=True)
    assert n == 70 and host.total_count == 70           # BASELINE config 3: every self-attention layer
    uninstall(host)
    assert host.SpatialAttnProcessor2_0 is object and host.cal_attn_mask_xl is rp.cal_attn_mask_xl


def test_processor_is_deepcopy_and_module_safe():
    p = spider_b200.SpatialAttnProcessor2_0(id_length=3)
    q = copy.deepcopy(p)
    assert q.id_length == 3 and q.total_length == 4 and isinstance(q.id_bank, IdBank) and q.id_bank is not p.id_bank
    assert q._host is p._host
    a = FakeAttention(128, 2)
    a.set_processor(q)
    a.to(torch.float16)
    assert a.processor is q and list(q.parameters()) == []


# ------------------------------------------------------------------------------------------ batched issue (8f.4)
class _FakeLib:
    """Records what reaches the library: direct entry points and the contents of csa_run_batch."""

    def __init__(self):
        self.log = []

    def csa_run_batch(self, arr, n, stream, failed):
        self.log.append(("batch", [arr[i].kind for i in range(n)], stream))
        return 0

    def csa_linear(self, a, stream):
        self.log.append(("linear", stream))
        return 0

    def csa_gather_rows(self, *a):
        self.log.append(("gather_rows",))
        return 0

    def csa_last_error(self):
        return b""


def test_batch_defers_in_order_and_flushes_before_immediate_calls(monkeypatch):
    lib = _FakeLib()
    monkeypatch.setattr(native, "load", lambda: lib)
    monkeypatch.setattr(native, "_stream_ptr", lambda t: 7)
    monkeypatch.setattr(native, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(native, "ensure_device", lambda d: None)
    monkeypatch.setattr(native, "_linear_workspace", lambda d, st: torch.zeros(64, dtype=torch.uint8))
    monkeypatch.setattr(native, "_on_device_of", lambda t: native._NULL_CTX)
    x = torch.zeros((8, 16), dtype=torch.bfloat16)
    w = torch.zeros((32, 16), dtype=torch.bfloat16)
    # no batch open: immediate
    native.linear(x, w)
    assert lib.log == [("linear", 7)]
    lib.log.clear()
    # open batch: nothing reaches the library until the flush, then everything in order in ONE call
    native.begin_batch(x)
    y = native.linear(x, w)
    native.linear(y, torch.zeros((16, 32), dtype=torch.bfloat16), torch.zeros(16, dtype=torch.bfloat16))
    assert lib.log == []
    native.flush_batch()
    assert lib.log == [("batch", [native.CSA_CALL_LINEAR, native.CSA_CALL_LINEAR], 7)]
    native.flush_batch()                      # closed: no-op
    assert len(lib.log) == 1
    lib.log.clear()
    # a wrapper that launches immediately flushes what was deferred first and closes the batch
    native.begin_batch(x)
    native.linear(x, w)
    idx = torch.zeros(4, dtype=torch.int32)
    native.gather_rows(x, idx, 4)
    native.linear(x, w)                        # after the flush: immediate again
    assert lib.log == [("batch", [native.CSA_CALL_LINEAR], 7), ("gather_rows",), ("linear", 7)]
    lib.log.clear()
    # error paths drop the batch
    native.begin_batch(x)
    native.linear(x, w)
    native.abort_batch()
    native.flush_batch()
    assert lib.log == []
