"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI, against
  * the golden fixtures produced by the unmodified reference (tests/golden/*.npz),
  * the CPU oracle (oracle/) on seeded inputs at sizes it finishes in seconds,
  * size-independent properties at BASELINE.json's full sizes.
Index lists are compared bit-exactly; attention outputs within the tolerance BASELINE.json states for this path:
max-abs <= 2e-2 and cosine >= 0.9995 against the reference's fp32 output.
"""
import copy
import random

import numpy as np
import pytest
import torch

import spider_b200
from spider_b200 import masks as csa_masks
from spider_b200 import native
from spider_b200.install import make_processor_class
from oracle import c_oracle
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention

from helpers import MAX_ABS, MIN_COS, attn_from_fixture, load_npz, max_abs_cos, unpack_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _assert_close(got, want, what):
    err, cos = max_abs_cos(got, want)
    assert err <= MAX_ABS and cos >= MIN_COS, f"{what}: max-abs {err:.3e} cos {cos:.6f}"
    return err, cos


def test_native_library_is_the_loaded_path():
    lib = native.load()
    assert lib.csa_abi_version() == native.CSA_ABI_VERSION
    native.ensure_device(torch.device(DEV))
    with open("/proc/self/maps") as f:
        assert "libcsa_b200.so" in f.read()


# ------------------------------------------------------------------------------------------------ compaction
def test_compaction_bit_exact_vs_reference_golden():
    z = load_npz("masks.npz")
    for ci in range(int(z["n_cases"])):
        seed, T, Fl, h, w = (int(x) for x in z[f"c{ci}_params"])
        for tag in ("32", "16"):
            n = int(z[f"c{ci}_n{tag}"])
            rows = unpack_rows(z[f"c{ci}_rows{tag}"], T * n)
            # (a) from the sample vector (what the processor does for regenerated masks)
            sample = rows[T - 1].copy()
            sample[Fl * n:] = False       # any values there are discarded by the mask algebra
            sample[Fl * n:] = np.random.RandomState(ci).rand(n) < 0.5   # ... so put noise to prove it
            r = torch.from_numpy(sample).to(DEV)
            idx, counts = native.compact_rows(r, T, T * n, 0, block_n=n, limit_cols=Fl * n)
            # (b) from the rows of a dense mask (what the processor does with the driver's first mask)
            dense = torch.from_numpy(rows).to(DEV)
            idx2, counts2 = native.compact_rows(dense, T, T * n, dense.stride(0))
            torch.cuda.synchronize()
            assert counts.cpu().numpy().tolist() == z[f"c{ci}_counts{tag}"].tolist()
            assert torch.equal(counts, counts2)
            for f in range(T):
                want = z[f"c{ci}_idx{tag}_{f}"]
                c = int(counts[f])
                assert np.array_equal(idx[f, :c].cpu().numpy(), want), f"case {ci} tag {tag} row {f} (sample path)"
                assert np.array_equal(idx2[f, :c].cpu().numpy(), want), f"case {ci} tag {tag} row {f} (dense path)"


@pytest.mark.parametrize("n_cols,p", [(1, 1.0), (1, 0.0), (15, 0.5), (16, 0.5), (17, 0.5), (16384, 0.5),
                                      (16385, 0.9), (20480, 0.0), (20480, 1.0), (5 * 9216, 0.5), (33 * 4096, 0.3)])
def test_compaction_edge_sizes_vs_c_oracle(n_cols, p):
    g = torch.Generator().manual_seed(n_cols)
    rows = torch.rand((3, n_cols), generator=g) < p
    d = rows.to(DEV)
    idx, counts = native.compact_rows(d, 3, n_cols, d.stride(0))
    torch.cuda.synchronize()
    for r in range(3):
        want = c_oracle.nonzero(rows[r].numpy())
        assert int(counts[r]) == want.size
        assert np.array_equal(idx[r, :want.size].cpu().numpy(), want)
    # unaligned row starts (stride not a multiple of 16) take the byte path
    if n_cols > 40:
        sub = d[:, 3:n_cols - 2]
        idx, counts = native.compact_rows(sub, 3, n_cols - 5, d.stride(0))
        torch.cuda.synchronize()
        for r in range(3):
            want = c_oracle.nonzero(rows[r, 3:n_cols - 2].numpy())
            assert int(counts[r]) == want.size and np.array_equal(idx[r, :want.size].cpu().numpy(), want)


def test_full_size_mask_pipeline_1024sq():
    """BASELINE size (1024^2, T=5): reference-layout dense mask (419 MB) -> validate -> compact == sample path."""
    T, Fl, H = 5, 4, 1024
    torch.manual_seed(0)
    cm32, cm16 = csa_masks.cal_attn_mask_xl(T, Fl, 0.5, 0.5, H, H, device=DEV, dtype=torch.float16)
    for cm in (cm32, cm16):
        idx, counts = cm.lists()
        dense = cm.dense()
        assert dense.shape == cm.shape
        assert int(native.validate_mask(dense, cm.n_tokens).item()) == 0
        cm2 = csa_masks.from_dense(dense, T, Fl, validate=True)
        idx2, counts2 = cm2.lists()
        assert torch.equal(counts, counts2)
        for f in range(T):
            c = int(counts[f])
            assert torch.equal(idx[f, :c], idx2[f, :c])
            assert torch.equal(idx[f, :c].long(), torch.nonzero(dense[f * cm.n_tokens])[:, 0])
        # sortedness + ownership properties
        for f in range(Fl):
            c = int(counts[f])
            l = idx[f, :c]
            assert bool((l[1:] > l[:-1]).all()) and int(l[-1]) < Fl * cm.n_tokens
            own = (l >= f * cm.n_tokens) & (l < (f + 1) * cm.n_tokens)
            assert int(own.sum()) == cm.n_tokens
        # corrupt one byte -> the validator must see it and from_dense must raise
        dense[cm.n_tokens + 3, 17] ^= True
        assert int(native.validate_mask(dense, cm.n_tokens).item()) == 1
        with pytest.raises(ValueError):
            csa_masks.from_dense(dense, T, Fl, validate=True)
        del dense


def test_gather_rows_bit_exact():
    g = torch.Generator().manual_seed(0)
    src = torch.randn((5000, 640), generator=g).to(DEV, torch.bfloat16)
    idx = torch.randperm(4000, generator=g)[:1500].sort().values.int().to(DEV)
    cnt = torch.tensor([1500], dtype=torch.int32, device=DEV)
    out = native.gather_rows(src, idx, 1500, row_base=1000, count=cnt)
    assert torch.equal(out, src[idx.long() + 1000])
    out2 = torch.zeros((1500, 640), dtype=torch.bfloat16, device=DEV)
    native.gather_rows(src, idx, 1500, row_base=0, count=cnt, count_adjust=-500, out=out2)
    assert torch.equal(out2[:1000], src[idx[:1000].long()]) and float(out2[1000:].abs().sum()) == 0.0


def test_sample_ranges_and_gather_kv_bit_exact():
    """csa_sample_ranges / csa_gather_kv against numpy: the runs of the shared sample list S each frame attends, and
    the sampled K/V rows made contiguous (zero-filled for one tile past the count)."""
    for (Fl, N, C, sa, seed) in [(4, 1024, 1280, 0.5, 0), (3, 100, 64, 0.3, 1), (4, 64, 128, 0.0, 2),
                                 (4, 64, 128, 1.0, 3), (16, 256, 640, 0.5, 4)]:
        g = torch.Generator().manual_seed(seed)
        T = Fl + 1
        sample = torch.rand((T * N,), generator=g) < sa
        cm = csa_masks.CompactMask(T, Fl, N, sample=sample.to(DEV))
        s_idx, s_count, ranges = cm.sample_list(DEV)
        S = np.nonzero(sample[:Fl * N].numpy())[0].astype(np.int32)
        cnt = int(s_count.item())
        assert cnt == S.size and np.array_equal(s_idx[:cnt].cpu().numpy(), S)
        want = []
        for f in range(Fl):
            lo = int(np.searchsorted(S, f * N))
            hi = int(np.searchsorted(S, (f + 1) * N))
            want.append([0, lo, hi, S.size - hi])
            # the two runs + the own block are exactly the oracle's key list of frame f
            keys = np.concatenate([S[:lo], np.arange(f * N, (f + 1) * N), S[hi:]])
            assert np.array_equal(np.sort(keys), rp.index_lists(rp.frame_rows(sample, T, Fl))[f].numpy()[
                rp.index_lists(rp.frame_rows(sample, T, Fl))[f].numpy() < Fl * N])
        want.append([0, S.size, 0, 0])
        assert np.array_equal(ranges.cpu().numpy(), np.array(want, dtype=np.int32))
        k = torch.randn((2 * Fl * N, C), generator=g).to(torch.bfloat16).to(DEV)
        v = torch.randn((2 * Fl * N, C), generator=g).to(torch.bfloat16).to(DEV)
        k_s, v_s, cap = native.gather_kv(k, v, Fl * N, 2, s_idx, s_count, Fl * N)
        assert cap == Fl * N + native.CSA_TILE
        Sl = torch.from_numpy(S).long().to(DEV)
        for gi in range(2):
            assert torch.equal(k_s[gi * cap:gi * cap + cnt], k[gi * Fl * N + Sl])
            assert torch.equal(v_s[gi * cap:gi * cap + cnt], v[gi * Fl * N + Sl])
            pad = min(cap, cnt + native.CSA_TILE)
            assert float(k_s[gi * cap + cnt:gi * cap + pad].float().abs().sum()) == 0.0
            assert float(v_s[gi * cap + cnt:gi * cap + pad].float().abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------ attention kernel
def _kernel_vs_oracle(Fl, N, C, heads, sa, dtype, mode, seed=0):
    """native.attn_fwd against oracle.gathered_attention (CPU fp32) on the same 16-bit inputs."""
    g = torch.Generator().manual_seed(seed)
    T = Fl + 1
    sample = torch.rand((T * N,), generator=g) < sa
    rows = rp.frame_rows(sample, T, Fl)
    lists = rp.index_lists(rows)
    r = sample.to(DEV)
    idx, counts = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=Fl * N)

    def mk(rows_):
        return torch.randn((rows_, C), generator=g).to(dtype)

    pre = mode.endswith("_pre")     # sampled rows gathered once (csa_gather_kv) instead of in-kernel gather4
    mode = mode[:-4] if pre else mode
    if pre:
        s_idx, s_count, ranges = csa_masks.CompactMask(T, Fl, N, sample=r).sample_list(DEV)
    if mode == "write":
        q, k, v = mk(2 * Fl * N), mk(2 * Fl * N), mk(2 * Fl * N)
        qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
        o = torch.empty_like(qd)
        if pre:
            k_s, v_s, cap = native.gather_kv(kd, vd, Fl * N, 2, s_idx, s_count, Fl * N)
            native.attn_fwd(qd, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k_s, v_a=v_s, a_group_rows=cap,
                            ranges=ranges, range_base=0, range_step=1, k_b=kd, v_b=vd, b_group_rows=Fl * N,
                            cb=(0, N, N))
        else:
            native.attn_fwd(qd, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=kd, v_a=vd,
                            a_group_rows=Fl * N, idx=idx, counts=counts, list_base=0, list_step=1)
        want = rp.gathered_attention(q.view(2, Fl * N, C), k.view(2, Fl * N, C), v.view(2, Fl * N, C), lists[:Fl],
                                     heads).reshape(2 * Fl * N, C)
    elif mode in ("read", "read_early"):
        q, kc, vc = mk(2 * N), mk(2 * N), mk(2 * N)
        kb, vb = mk(2 * Fl * N), mk(2 * Fl * N)
        qd = q.to(DEV)
        o = torch.empty_like(qd)
        kw = dict(heads=heads, n_groups=2, n_frames=1, n_q=N, k_a=kb.to(DEV), v_a=vb.to(DEV), a_group_rows=Fl * N,
                  k_b=kc.to(DEV), v_b=vc.to(DEV), b_group_rows=N, cb=(0, 0, N))
        if mode == "read" and pre:
            k_s, v_s, cap = native.gather_kv(kw["k_a"], kw["v_a"], Fl * N, 2, s_idx, s_count, Fl * N)
            kw.update(k_a=k_s, v_a=v_s, a_group_rows=cap)
            native.attn_fwd(qd, o, ranges=ranges, range_base=Fl, range_step=0, **kw)
            keys = lists[Fl]
        elif mode == "read":
            native.attn_fwd(qd, o, idx=idx, counts=counts, list_base=Fl, list_step=0, g_adjust=-N, **kw)
            keys = lists[Fl]
        else:
            native.attn_fwd(qd, o, ca=(0, 0, Fl * N), **kw)
            keys = torch.arange(T * N, dtype=torch.int32)
        # reference key order (Comic_Generation.py:92,162): [bank frames..., current frame] per CFG half
        kall = torch.cat((kb.view(2, Fl * N, C), kc.view(2, N, C)), dim=1)
        vall = torch.cat((vb.view(2, Fl * N, C), vc.view(2, N, C)), dim=1)
        want = rp.gathered_attention(q.view(2, N, C), kall, vall, [keys], heads).reshape(2 * N, C)
    else:  # standard
        B = 2 * Fl
        q, k, v = mk(B * N), mk(B * N), mk(B * N)
        qd = q.to(DEV)
        o = torch.empty_like(qd)
        native.attn_fwd(qd, o, heads=heads, n_groups=1, n_frames=B, n_q=N, k_b=k.to(DEV), v_b=v.to(DEV),
                        b_group_rows=B * N, cb=(0, N, N))
        ident = [torch.arange(N, dtype=torch.int32)]
        want = torch.cat([rp.gathered_attention(q.view(B, N, C)[b:b + 1], k.view(B, N, C)[b:b + 1],
                                                v.view(B, N, C)[b:b + 1], ident, heads) for b in range(B)])
        want = want.reshape(B * N, C)
    torch.cuda.synchronize()
    assert o.dtype == dtype and torch.isfinite(o.float()).all()
    return _assert_close(o, want, f"{mode} F={Fl} N={N} C={C} sa={sa} {dtype}")


@pytest.mark.parametrize("mode", ["write", "write_pre", "read", "read_pre", "read_early", "standard"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attention_modes_vs_oracle(mode, dtype):
    _kernel_vs_oracle(4, 256, 128, 2, 0.5, dtype, mode)


@pytest.mark.parametrize("Fl,N,C,heads,sa", [
    (4, 576, 128, 2, 0.5),     # 768^2 /32 layer: N not a multiple of 128 (ragged Q and K tiles)
    (3, 100, 64, 1, 0.3),      # tiny ragged, predict.py-style 3 identity frames
    (1, 128, 64, 1, 0.5),      # single frame
    (2, 2304, 64, 1, 0.5),     # 768^2 /16 layer
    (4, 256, 640, 10, 0.0),    # sa = 0: pure self-attention through the gather path
    (4, 256, 640, 10, 1.0),    # sa = 1: dense over all frames, every tile a contiguous run
    (8, 128, 128, 2, 0.05),    # sparse sample: gathered tiles only
    (4, 1024, 1280, 20, 0.5),  # the 32x32 SDXL layer of BASELINE config 2
])
@pytest.mark.parametrize("gather", ["inline", "pre"])
def test_attention_write_shapes_vs_oracle(Fl, N, C, heads, sa, gather):
    _kernel_vs_oracle(Fl, N, C, heads, sa, torch.bfloat16, "write" if gather == "inline" else "write_pre",
                      seed=N + Fl)


@pytest.mark.parametrize("Fl,N,C,heads,sa", [(4, 576, 128, 2, 0.5), (3, 100, 64, 1, 0.3), (4, 1024, 1280, 20, 0.5),
                                             (4, 256, 128, 2, 0.0), (4, 256, 128, 2, 1.0)])
@pytest.mark.parametrize("gather", ["inline", "pre"])
def test_attention_read_shapes_vs_oracle(Fl, N, C, heads, sa, gather):
    _kernel_vs_oracle(Fl, N, C, heads, sa, torch.float16, "read" if gather == "inline" else "read_pre", seed=N)


def test_attention_strided_inputs():
    """K/V/Q/O that are column slices of wider matrices (row stride > heads*64), as fused-QKV callers produce."""
    g = torch.Generator().manual_seed(2)
    N, C, heads = 256, 128, 2
    big = torch.randn((2 * N, 3 * C), generator=g).to(torch.bfloat16)
    bd = big.to(DEV)
    q, k, v = bd[:, :C], bd[:, C:2 * C], bd[:, 2 * C:]
    obig = torch.zeros((2 * N, 2 * C), dtype=torch.bfloat16, device=DEV)
    o = obig[:, C:]
    native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=2, n_q=N, k_b=k, v_b=v, b_group_rows=2 * N, cb=(0, N, N))
    ident = [torch.arange(N, dtype=torch.int32)]
    want = torch.cat([rp.gathered_attention(big[b * N:(b + 1) * N, :C][None], big[b * N:(b + 1) * N, C:2 * C][None],
                                            big[b * N:(b + 1) * N, 2 * C:][None], ident, heads) for b in range(2)])
    _assert_close(o, want.reshape(2 * N, C), "strided")
    assert float(obig[:, :C].abs().sum()) == 0.0


@pytest.mark.parametrize("Fl,N,C,heads,max_ctas,split", [
    (4, 1024, 128, 2, 0, True),      # 64 units on 148 SMs: every unit is split
    (4, 1024, 1280, 20, 0, True),    # the 32x32 SDXL layer: 640 units = 4 whole rounds + 48 split units
    (2, 512, 192, 3, 10, True),      # 10 CTAs, 24 units: 20 whole + 4 split in two
    (3, 200, 64, 1, 4, 3),           # ragged tiles, forced 3 pieces of 1-2 key tiles
    (1, 100, 64, 1, 0, 8),           # 2 units of 1-2 key tiles forced into 8 pieces: most pieces are empty
])
def test_attention_tail_split_matches_whole_units(Fl, N, C, heads, max_ctas, split):
    """Units of the last, partial round are cut into pieces along the keys and merged by the CTA that delivers the last
    piece: same result as processing every unit whole (fp32 rounding of the merge aside), launch after launch (the
    arrival counters are left zero), and within the parity bar of the oracle."""
    g = torch.Generator().manual_seed(N + heads)
    T = Fl + 1
    sample = torch.rand((T * N,), generator=g) < 0.5
    lists = rp.index_lists(rp.frame_rows(sample, T, Fl))
    r = sample.to(DEV)
    s_idx, s_count, ranges = csa_masks.CompactMask(T, Fl, N, sample=r).sample_list(DEV)
    q, k, v = (torch.randn((2 * Fl * N, C), generator=g).to(torch.bfloat16) for _ in range(3))
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    k_s, v_s, cap = native.gather_kv(kd, vd, Fl * N, 2, s_idx, s_count, Fl * N)
    kw = dict(heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=ranges,
              range_base=0, range_step=1, k_b=kd, v_b=vd, b_group_rows=Fl * N, cb=(0, N, N), max_ctas=max_ctas)
    whole = native.attn_fwd(qd, torch.empty_like(qd), split=False, **kw)
    assert native.last_launch()["split"] == 1
    outs = [native.attn_fwd(qd, torch.full_like(qd, float("nan")), split=split, **kw) for _ in range(3)]
    assert native.last_launch()["split"] > 1, native.last_launch()
    torch.cuda.synchronize()
    for o in outs:
        assert torch.isfinite(o.float()).all()
        assert torch.equal(o, outs[0])                      # deterministic, counters reset between launches
        assert (o.float() - whole.float()).abs().max().item() <= 4e-3
    hdr = native.attn_workspace(qd.device)[:4096]
    assert int(hdr.to(torch.int32).sum()) == 0
    want = rp.gathered_attention(q.view(2, Fl * N, C), k.view(2, Fl * N, C), v.view(2, Fl * N, C), lists[:Fl],
                                 heads).reshape(2 * Fl * N, C)
    _assert_close(outs[0], want, f"tail split F={Fl} N={N} C={C}")


# ------------------------------------------------------------------------------------------------ processor vs reference
def _gpu_attn(z, prefix, C, heads, dtype):
    return attn_from_fixture(z, prefix, C, heads, dtype=dtype, device=DEV)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_processor_calls_vs_reference_golden(dtype):
    """calls.npz: the reference's __call1__/__call2__ fp32 outputs for a supplied dense mask, write and read."""
    z = load_npz("calls.npz")
    H, W, Fl, C, heads = (int(x) for x in z["geom"])
    n = (H // 16) * (W // 16)
    rows = torch.from_numpy(unpack_rows(z["rows16"], (Fl + 1) * n))
    dense = rp.dense_mask(rows).to(DEV)
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 1
    host.mask4096 = dense
    host.mask1024 = torch.zeros((5, 5), dtype=torch.bool, device=DEV)
    cls = make_processor_class(host)
    attn = _gpu_attn(z, "attn_", C, heads, dtype)
    hs_w = torch.from_numpy(z["hs_w"]).to(DEV, dtype)
    hs_r = torch.from_numpy(z["hs_r"]).to(DEV, dtype)

    def fresh():
        host.mask4096, host.attn_count = dense, 0
        return cls(id_length=Fl, device=DEV, dtype=torch.float16)

    with torch.no_grad():
        p = fresh()
        host.write, host.cur_step = True, 0
        out = p(attn, hs_w)                                   # early step -> standard
        _assert_close(out, torch.from_numpy(z["write_standard"]), "write_standard")
        assert out.shape == hs_w.shape and out.dtype == dtype and out.device == hs_w.device
        # gate forced to the consistent branch: cur_step 25 and a draw > 0.1
        p = fresh()
        host.write, host.cur_step = True, 25
        random.seed(0)   # 0.844
        out = p(attn, hs_w)
        assert p._last_branch == "consistent"
        _assert_close(out, torch.from_numpy(z["write_consistent"]), "write_consistent")
        # read through the bank written at the same step
        host.mask4096, host.write, host.cur_step = dense, False, 25
        random.seed(0)
        out = p(attn, hs_r)
        assert p._last_branch == "consistent"
        _assert_close(out, torch.from_numpy(z["read_consistent"]), "read_consistent")
        # read-early: write at step 0, read at step 0
        p = fresh()
        host.write, host.cur_step = True, 0
        p(attn, hs_w)
        host.write, host.cur_step = False, 0
        out = p(attn, hs_r)
        _assert_close(out, torch.from_numpy(z["read_early"]), "read_early")


@pytest.mark.parametrize("bank_store,kv_gather", [("kv", "pre"), ("hidden", "pre"), ("both", "pre"), ("kv", "inline")])
def test_story_state_machine_vs_reference_golden(bank_store, kv_gather):
    """The whole Appendix-C scenario on the GPU: 8 write steps + 8 read steps x 3 layers, same seeds, same inputs,
    same gates, same sample vectors (processors draw them on the CPU generator like the golden run) — every call's
    output against the reference's."""
    z = load_npz("story.npz")
    H, W, Fl, C, heads, steps = (int(x) for x in z["geom"])
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count, host.sa32, host.sa64 = H, W, 3, 0.5, 0.5
    cls = make_processor_class(host, bank_store=bank_store)
    cls.kv_gather = kv_gather
    rp.setup_seed(2047)
    attns = [FakeAttention(C, heads) for _ in range(3)]            # consumes the torch stream like the golden run
    for li, a in enumerate(attns):
        for k_, v_ in a.state_dict().items():
            assert torch.equal(v_, torch.from_numpy(z[f"attn{li}_{k_}"]))
    attns = [a.to(DEV, torch.bfloat16) for a in attns]
    procs = copy.deepcopy([cls(id_length=Fl, device="cpu", dtype=torch.float32) for _ in range(3)])
    m32, m16 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)     # the driver's dense masks (:376)
    host.mask1024, host.mask4096 = m32.to(DEV), m16.to(DEV)
    draws, worst = [], (0.0, 1.0)
    orig = random.random
    try:
        random.random = lambda: (draws.append(orig()), draws[-1])[1]
        with torch.no_grad():
            for phase, write in (("w", True), ("r", False)):
                host.write, host.cur_step = write, 0
                for s in range(steps):
                    for li, p in enumerate(procs):
                        x = torch.from_numpy(z[f"{phase}{s}_{li}_in"]).to(DEV, torch.bfloat16)
                        out = p(attns[li], x)
                        e, c = _assert_close(out, torch.from_numpy(z[f"{phase}{s}_{li}_out"]), f"{phase}{s}_{li}")
                        worst = (max(worst[0], e), min(worst[1], c))
    finally:
        random.random = orig
    assert np.array_equal(np.array(draws), z["draws"])
    assert host.cur_step == int(z["final_cur_step"])
    for p, keys in zip(procs, z["bank_keys"]):
        assert sorted(p.id_bank.keys()) == list(keys)
    print(f"story parity ({bank_store}, {kv_gather}): worst max-abs {worst[0]:.3e}, worst cos {worst[1]:.6f}")


def test_processor_vs_oracle_config2_sizes():
    """BASELINE config 2: a 32x32 (1280 ch, 20 heads) layer at 1024^2, id_length 4, write then read via the bank,
    bf16 on the GPU vs the CPU oracle processor in fp32 on identical inputs, weights and supplied mask."""
    H = W = 1024
    Fl, C, heads = 4, 1280, 20
    N = (H // 32) * (W // 32)
    torch.manual_seed(0)
    random.seed(0)
    attn = FakeAttention(C, heads)
    hs_w = torch.randn(2 * Fl, N, C)
    hs_r = torch.randn(2, N, C)
    m32, m16 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)
    del m16
    # oracle
    st = rp.StoryState(total_count=10 ** 9, height=H, width=W)
    st.mask1024 = m32
    orc = rp.ConsistentAttnOracle(st, id_length=Fl)
    with torch.no_grad():
        st.write, st.cur_step = True, 25
        random.seed(0)
        want_w = orc(attn, hs_w)
        st.write = False
        random.seed(0)
        want_r = orc(attn, hs_r)
    assert [t[0] for t in st.trace] == ["consistent", "consistent"]
    # product
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 10 ** 9
    host.mask1024 = m32.to(DEV)
    cls = make_processor_class(host)
    p = cls(id_length=Fl)
    ga = copy.deepcopy(attn).to(DEV, torch.bfloat16)
    with torch.no_grad():
        host.write, host.cur_step = True, 25
        random.seed(0)
        got_w = p(ga, hs_w.to(DEV, torch.bfloat16))
        host.write = False
        random.seed(0)
        got_r = p(ga, hs_r.to(DEV, torch.bfloat16))
    _assert_close(got_w, want_w, "config2 write 32x32")
    _assert_close(got_r, want_r, "config2 read 32x32")


# ------------------------------------------------------------------------------------------------ full-size properties
def _full_size_setup(Fl, N, C, sa, dtype, seed=0):
    torch.manual_seed(seed)
    T = Fl + 1
    r = torch.rand((T * N,), device=DEV) < sa
    idx, counts = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=Fl * N)
    q = torch.randn((2 * Fl * N, C), device=DEV, dtype=dtype)
    k = torch.randn((2 * Fl * N, C), device=DEV, dtype=dtype)
    v = torch.randn((2 * Fl * N, C), device=DEV, dtype=dtype)
    return r, idx, counts, q, k, v


def _write(q, k, v, idx, counts, Fl, N, heads):
    o = torch.empty_like(q)
    native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k, v_a=v, a_group_rows=Fl * N,
                    idx=idx, counts=counts, list_base=0, list_step=1)
    return o


def _torch_ref_frame(q, k, v, keys, heads):
    d = 64
    qh = q.float().view(-1, heads, d).transpose(0, 1)[None]
    kh = k[keys].float().view(-1, heads, d).transpose(0, 1)[None]
    vh = v[keys].float().view(-1, heads, d).transpose(0, 1)[None]
    o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)[0]
    return o.transpose(0, 1).reshape(q.shape[0], heads * d)


@pytest.mark.parametrize("Fl,N,C,heads", [(4, 4096, 640, 10), (4, 1024, 1280, 20), (16, 4096, 640, 10)])
def test_full_size_properties(Fl, N, C, heads):
    """BASELINE shapes (64x64 and 32x32 SDXL layers at 1024^2; F=16 where the reference cannot even build its mask):
    rows of P sum to 1, linearity in V, invariance to a key permutation, and a plain-torch fp32 gathered reference
    on sampled frames."""
    dtype = torch.bfloat16
    r, idx, counts, q, k, v = _full_size_setup(Fl, N, C, 0.5, dtype)
    # (1) V == 1  ->  O == 1 (softmax rows sum to one, whatever the key list)
    ones = torch.ones_like(v)
    o1 = _write(q, k, ones, idx, counts, Fl, N, heads)
    assert float((o1.float() - 1.0).abs().max()) <= 1e-2
    # (2) plain torch fp32 on two (group, frame) units
    o = _write(q, k, v, idx, counts, Fl, N, heads)
    for g_, f in ((0, 0), (1, Fl - 1)):
        c = int(counts[f])
        keys = idx[f, :c].long() + g_ * Fl * N
        qs = slice((g_ * Fl + f) * N, (g_ * Fl + f + 1) * N)
        _assert_close(o[qs], _torch_ref_frame(q[qs], k, v, keys, heads), f"full-size unit g{g_} f{f}")
    # (3) linearity in V
    v2 = torch.randn_like(v)
    o2 = _write(q, k, v2, idx, counts, Fl, N, heads)
    o12 = _write(q, k, (v.float() + v2.float()).to(dtype), idx, counts, Fl, N, heads)
    assert float((o12.float() - (o.float() + o2.float())).abs().max()) <= 3e-2
    # (4) reversing the token order inside every frame (keys, values and the sample vector alike) permutes nothing
    #     but the order in which keys are visited: outputs must agree up to rounding
    perm = torch.arange(N - 1, -1, -1, device=DEV)
    T = Fl + 1
    r_p = r.view(T, N)[:, perm].reshape(-1).contiguous()
    idx_p, counts_p = native.compact_rows(r_p, T, T * N, 0, block_n=N, limit_cols=Fl * N)
    assert torch.equal(counts_p[:Fl], counts[:Fl])
    k_p = k.view(2 * Fl, N, C)[:, perm].reshape(-1, C).contiguous()
    v_p = v.view(2 * Fl, N, C)[:, perm].reshape(-1, C).contiguous()
    o_p = _write(q, k_p, v_p, idx_p, counts_p, Fl, N, heads)
    assert float((o_p.float() - o.float()).abs().max()) <= 1e-2


def test_sa_one_equals_dense_and_sa_zero_equals_standard():
    Fl, N, C, heads = 4, 1024, 640, 10
    for sa in (0.0, 1.0):
        r, idx, counts, q, k, v = _full_size_setup(Fl, N, C, sa, torch.float16, seed=3)
        o = _write(q, k, v, idx, counts, Fl, N, heads)
        o_ref = torch.empty_like(q)
        if sa == 0.0:   # every frame attends only itself == the standard branch
            native.attn_fwd(q, o_ref, heads=heads, n_groups=1, n_frames=2 * Fl, n_q=N, k_b=k, v_b=v,
                            b_group_rows=2 * Fl * N, cb=(0, N, N))
        else:           # every frame attends every key of its CFG half == contiguous segment over the group
            native.attn_fwd(q, o_ref, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k, v_a=v,
                            a_group_rows=Fl * N, ca=(0, 0, Fl * N))
        assert float((o.float() - o_ref.float()).abs().max()) <= 2e-3


@pytest.mark.parametrize("cur_step,kv_gather", [(2, "pre"), (25, "pre"), (25, "inline")])
def test_batched_read_equals_separate_reads(cur_step, kv_gather):
    """Opt-in batched read pass: R generated frames in one call (batch 2*R, [uncond R, cond R]) give,
    frame by frame, the output of R reference-style batch-2 calls — early (all bank keys) and consistent branch."""
    H = W = 512
    Fl, C, heads, R = 4, 128, 2, 3
    N = (H // 32) * (W // 32)
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 10 ** 9
    cls = make_processor_class(host)
    cls.kv_gather = kv_gather
    torch.manual_seed(5)
    attn = FakeAttention(C, heads).to(DEV, torch.float16)
    host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device=DEV)
    proc = cls(id_length=Fl, device=DEV)
    xw = torch.randn(2 * Fl, N, C, device=DEV, dtype=torch.float16)
    xr = torch.randn(2, R, N, C, device=DEV, dtype=torch.float16)      # [cfg half][frame]
    real = random.random
    random.random = lambda: 0.99
    try:
        with torch.no_grad():
            host.write, host.cur_step = True, cur_step
            proc(attn, xw)
            host.write = False
            singles = []
            for r in range(R):
                host.cur_step = cur_step
                singles.append(proc(attn, xr[:, r].contiguous()))      # (2, N, C)
            with pytest.raises(ValueError):
                proc(attn, xr.reshape(2 * R, N, C))                    # batch 2*R is refused unless opted in
            proc.batched_read = True
            host.cur_step = cur_step
            got = proc(attn, xr.reshape(2 * R, N, C)).view(2, R, N, C)
    finally:
        random.random = real
    assert proc._last_branch == ("early" if cur_step < 5 else "consistent")
    for r in range(R):
        assert torch.equal(got[:, r], singles[r]), f"frame {r}"
