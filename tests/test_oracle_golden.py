"""CPU: the oracle (oracle/reference_port.py, oracle/compact_ref.c) against the golden fixtures produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import random

import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_loader
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention

from helpers import attn_from_fixture, load_npz, unpack_rows


def _mask_cases():
    z = load_npz("masks.npz")
    for ci in range(int(z["n_cases"])):
        seed, T, Fl, h, w = (int(x) for x in z[f"c{ci}_params"])
        sa32, sa64 = (float(x) for x in z[f"c{ci}_sa"])
        yield z, ci, seed, T, Fl, h, w, sa32, sa64, getattr(torch, str(z[f"c{ci}_dtype"]))


def test_mask_sampler_matches_reference_rng_and_rows():
    """Same torch seed -> same two sample vectors -> same distinct rows as the reference's cal_attn_mask_xl."""
    for z, ci, seed, T, Fl, h, w, sa32, sa64, dt in _mask_cases():
        torch.manual_seed(seed)
        r32, r16 = rp.sample_vectors(T, sa32, sa64, h, w, "cpu", dt)
        for tag, r in (("32", r32), ("16", r16)):
            n = int(z[f"c{ci}_n{tag}"])
            want = unpack_rows(z[f"c{ci}_rows{tag}"], T * n)
            got = rp.frame_rows(r, T, Fl).numpy()
            assert got.shape == want.shape
            assert (got == want).all(), f"case {ci} rows{tag}"
            assert (got.sum(1) == z[f"c{ci}_counts{tag}"]).all()


def test_dense_mask_equals_reference_layout():
    for z, ci, seed, T, Fl, h, w, sa32, sa64, dt in _mask_cases():
        if h * w > 256 * 256:
            continue
        torch.manual_seed(seed)
        m32, m16 = rp.cal_attn_mask_xl(T, Fl, sa32, sa64, h, w, "cpu", dt)
        for tag, m in (("32", m32), ("16", m16)):
            n = int(z[f"c{ci}_n{tag}"])
            want = unpack_rows(z[f"c{ci}_rows{tag}"], T * n)
            assert m.shape == (T * n, T * n) and m.dtype == torch.bool
            assert (m[::n].numpy() == want).all()
            assert c_oracle.blocks_uniform(m.numpy(), n)


def test_index_lists_python_and_c_oracles_match_golden():
    for z, ci, seed, T, Fl, h, w, sa32, sa64, dt in _mask_cases():
        for tag in ("32", "16"):
            n = int(z[f"c{ci}_n{tag}"])
            rows = unpack_rows(z[f"c{ci}_rows{tag}"], T * n)
            lists = rp.index_lists(torch.from_numpy(rows))
            # sample vector is not stored; rebuild the C rows from a vector that reproduces them: row F restricted
            # to the id columns is exactly the sample restricted to those columns
            sample = rows[T - 1].copy()
            sample[Fl * n:] = False
            for f in range(T):
                want = z[f"c{ci}_idx{tag}_{f}"]
                assert np.array_equal(lists[f].numpy(), want)
                assert np.array_equal(c_oracle.nonzero(rows[f]), want)
                crow = c_oracle.frame_row(sample, T, Fl, f).astype(bool)
                assert np.array_equal(crow, rows[f]), f"C frame_row case {ci} tag {tag} row {f}"


def test_direct_calls_match_reference_outputs():
    z = load_npz("calls.npz")
    H, W, Fl, C, heads = (int(x) for x in z["geom"])
    attn = attn_from_fixture(z, "attn_", C, heads)
    n = (H // 16) * (W // 16)
    rows = torch.from_numpy(unpack_rows(z["rows16"], (Fl + 1) * n))
    mask = rp.dense_mask(rows)
    cut = Fl * n
    st = rp.StoryState()
    proc = rp.ConsistentAttnOracle(st, id_length=Fl)
    hs_w, hs_r = torch.from_numpy(z["hs_w"]), torch.from_numpy(z["hs_r"])
    with torch.no_grad():
        got = {
            "write_consistent": proc.consistent(attn, hs_w, None, mask[:cut, :cut]),
            "write_standard": proc.standard(attn, hs_w, None, None),
        }
        enc = torch.cat((hs_w[:Fl], hs_r[:1], hs_w[Fl:], hs_r[1:]))
        got["read_consistent"] = proc.consistent(attn, hs_r, enc, mask[cut:])
        got["read_early"] = proc.standard(attn, hs_r, enc, None)
    for k, v in got.items():
        assert torch.allclose(v, torch.from_numpy(z[k]), atol=1e-5, rtol=1e-5), k


def test_gathered_form_equals_dense_masked_form():
    """The per-frame index-list formulation the CUDA path implements == the reference's dense-mask SDPA."""
    z = load_npz("calls.npz")
    H, W, Fl, C, heads = (int(x) for x in z["geom"])
    attn = attn_from_fixture(z, "attn_", C, heads)
    n = (H // 16) * (W // 16)
    rows = torch.from_numpy(unpack_rows(z["rows16"], (Fl + 1) * n))
    lists = rp.index_lists(rows)
    hs_w = torch.from_numpy(z["hs_w"])
    with torch.no_grad():
        x = hs_w.view(2, Fl * n, C)
        q, k, v = attn.to_q(x), attn.to_k(x), attn.to_v(x)
        o = rp.gathered_attention(q, k, v, lists[:Fl], heads).reshape(2 * Fl, n, C)
        out = attn.to_out[0](o)
    assert torch.allclose(out, torch.from_numpy(z["write_consistent"]), atol=2e-5, rtol=1e-4)
    # independent C restatement (double accumulation), one head of one frame
    f, hd = 2, 1
    qh = q[1, f * n:(f + 1) * n, hd * 64:(hd + 1) * 64].numpy()
    kh = k[1, :, hd * 64:(hd + 1) * 64].numpy()
    vh = v[1, :, hd * 64:(hd + 1) * 64].numpy()
    c_out = c_oracle.attention_f32(qh, kh, vh, lists[f].numpy(), 64 ** -0.5)
    assert np.allclose(c_out, o.view(2, Fl * n, C)[1, f * n:(f + 1) * n, hd * 64:(hd + 1) * 64].numpy(), atol=2e-5)


def _run_story_with_oracle(z):
    H, W, Fl, C, heads, steps = (int(x) for x in z["geom"])
    rp.setup_seed(2047)
    attns = [FakeAttention(C, heads) for _ in range(3)]   # consumes the torch stream like the golden run
    for li, a in enumerate(attns):   # default init under the same seed must reproduce the stored weights
        for k1, v1 in a.state_dict().items():
            assert torch.equal(v1, torch.from_numpy(z[f"attn{li}_{k1}"])), "nn.Linear init drifted from golden run"
    st = rp.StoryState(total_count=3, sa32=0.5, sa64=0.5, height=H, width=W)
    procs = [rp.ConsistentAttnOracle(st, id_length=Fl) for _ in range(3)]
    st.mask1024, st.mask4096 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)
    outs = {}
    with torch.no_grad():
        for phase, write in (("w", True), ("r", False)):
            st.write, st.cur_step = write, 0
            for s in range(steps):
                for li, p in enumerate(procs):
                    x = torch.from_numpy(z[f"{phase}{s}_{li}_in"])
                    outs[f"{phase}{s}_{li}_out"] = p(attns[li], x)
    return st, procs, outs


def test_story_state_machine_matches_reference():
    """SURVEY.md Appendix C scenario: gates, branches, bank keys, step counter and every call's output."""
    z = load_npz("story.npz")
    st, procs, outs = _run_story_with_oracle(z)
    draws = [t[2] for t in st.trace if t[2] is not None]
    assert np.allclose(draws, z["draws"], atol=0, rtol=0)
    assert [round(d, 4) for d in draws[:6]] == [0.6338, 0.4726, 0.4165, 0.4051, 0.8172, 0.8005]
    want_trace = [str(t) for t in z["trace"]]
    got_trace = []
    for kind, step, u in st.trace:
        got_trace.append("consistent" if kind == "consistent" else "standard")
    assert got_trace == [t.split(":")[0] for t in want_trace]
    assert [t[1] for t in st.trace] == [int(t.split(":")[1]) for t in want_trace]
    assert st.cur_step == int(z["final_cur_step"])
    for p, keys in zip(procs, z["bank_keys"]):
        assert sorted(p.id_bank.keys()) == list(keys)
    for k, v in outs.items():
        assert torch.allclose(v, torch.from_numpy(z[k]), atol=1e-5, rtol=1e-5), k


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_against_live_reference():
    """When the reference tree is present, run it live next to the port on a fresh random case."""
    ref = ref_loader.load_reference()
    H = W = 96
    Fl, C, heads = 3, 128, 2
    n = (H // 16) * (W // 16)
    torch.manual_seed(5)
    import sys
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    m32_ref, m16_ref = gu.cal_attn_mask_xl(Fl + 1, Fl, 0.4, 0.6, H, W, device="cpu", dtype=torch.float32)
    torch.manual_seed(5)
    m32, m16 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.4, 0.6, H, W)
    assert torch.equal(m32, m32_ref) and torch.equal(m16, m16_ref)
    attn = FakeAttention(C, heads)
    hs = torch.randn(2 * Fl, n, C)
    rproc = ref.SpatialAttnProcessor2_0(id_length=Fl, device="cpu", dtype=torch.float32)
    oproc = rp.ConsistentAttnOracle(rp.StoryState(), id_length=Fl)
    cut = Fl * n
    with torch.no_grad():
        a = rproc.__call1__(attn, hs, None, m16_ref[:cut, :cut])
        b = oproc.consistent(attn, hs, None, m16[:cut, :cut])
    assert torch.allclose(a, b, atol=1e-6)
