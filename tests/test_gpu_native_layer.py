"""GPU tests of SURVEY 8f.4: the projections either side of the attention issued by the library (csa_linear, a plain
cuBLASLt GEMM with nn.Linear's layouts) and the one-call batch (csa_run_batch).  The reference is
attn.to_q / to_k / to_v / to_out[0] (StoryDiffusion/Comic_Generation.py:155,164-165,185), i.e. torch's own nn.Linear
on the same weights; a floating-point kernel, so the comparison is against torch fp32 with a stated tolerance."""
import random

import pytest
import torch

import spider_b200
from spider_b200 import native
from spider_b200.install import make_processor_class
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention

from helpers import MAX_ABS, MIN_COS, max_abs_cos

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k,bias", [(8192, 1280, 1280, False), (4096, 2560, 1280, False), (1000, 640, 640, True),
                                        (24, 1280, 640, True)])
def test_linear_matches_torch_fp32(dtype, m, n, k, bias):
    g = torch.Generator(device=DEV).manual_seed(m + n)
    x = torch.randn((m, k), device=DEV, generator=g).to(dtype)
    w = (torch.randn((n, k), device=DEV, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn((n,), device=DEV, generator=g).to(dtype) if bias else None
    want = torch.nn.functional.linear(x.float(), w.float(), None if b is None else b.float())
    got = native.linear(x, w, b)
    # fp32 accumulation, one rounding to 16 bits: |y| ~ 1, half an ulp of bf16 is 2^-9 relative
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert (got.float() - want).abs().max().item() <= tol
    # strided output (K and V land in the two halves of one buffer) and strided input rows
    buf = torch.zeros((m, 2 * n), dtype=dtype, device=DEV)
    native.linear(x, w, b, out=buf[:, n:])
    assert torch.equal(buf[:, n:], got) and float(buf[:, :n].abs().max()) == 0.0
    x2 = torch.zeros((m, k + 8), dtype=dtype, device=DEV)
    x2[:, :k] = x
    assert torch.equal(native.linear(x2[:, :k], w, b), got)


def test_linear_rejects_bad_arguments():
    x = torch.zeros((8, 64), dtype=torch.bfloat16, device=DEV)
    w = torch.zeros((16, 32), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(native.CsaNativeError):
        native.linear(x, w)
    with pytest.raises(native.CsaNativeError):
        native.linear(x, torch.zeros((16, 64), dtype=torch.float16, device=DEV))


def _story(native_projections, dtype=torch.bfloat16):
    """write (early, standard, consistent) + read (early, consistent) calls on one layer; returns outputs + trace."""
    H = W = 256
    Fl, C, heads = 4, 640, 10
    N = (H // 16) * (W // 16)
    torch.manual_seed(0)
    attn = FakeAttention(C, heads).to(DEV, dtype)
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 10 ** 9
    cls = make_processor_class(host)
    cls.native_projections = native_projections
    proc = cls(id_length=Fl)
    torch.manual_seed(3)
    host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device=DEV,
                                                                dtype=torch.float16)
    g = torch.Generator(device=DEV).manual_seed(11)
    xw = torch.randn((2 * Fl, N, C), device=DEV, generator=g).to(dtype)
    xr = torch.randn((2, N, C), device=DEV, generator=g).to(dtype)
    outs, trace = [], []
    real = random.random
    try:
        with torch.no_grad():
            for write, step, draw, x in ((True, 0, 0.0, xw), (True, 6, 0.05, xw), (True, 25, 0.9, xw),
                                         (False, 0, 0.0, xr), (False, 25, 0.9, xr)):
                host.write, host.cur_step = write, step
                random.random = lambda d=draw: d
                before = dict(native.LAUNCHES)
                outs.append(proc(attn, x).float())
                trace.append((proc._last_branch, _gemms(native.LAUNCHES) - _gemms(before)))
    finally:
        random.random = real
    torch.cuda.synchronize()
    return outs, trace


def _gemms(counters):
    """projections issued by the library: the hand-written GEMM (csa_gemm) or, for shapes it does not take, cuBLASLt"""
    return counters["csa_linear"] + counters["csa_gemm"]


def test_native_projections_match_module_projections():
    got, trace_n = _story(True)
    want, trace_m = _story(False)
    assert [t[0] for t in trace_n] == [t[0] for t in trace_m] == ["early", "standard", "consistent", "early",
                                                                    "consistent"]
    assert all(t[1] == 2 for t in trace_n), trace_n       # q|k|v (one stacked-weight GEMM), out: two GEMMs per call
    assert all(t[1] == 0 for t in trace_m), trace_m
    for a, b, t in zip(got, want, trace_n):
        err = (a - b).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0).item()
        # same weights, same kernels in between; only the GEMM algorithm may differ
        assert err <= 4e-3 and cos >= 0.99999, f"{t}: max-abs {err:.3e} cos {cos:.7f}"


def test_native_layer_matches_oracle():
    """The whole batched call (projections + gather + attention + output projection) against the CPU oracle."""
    H = W = 256
    Fl, C, heads = 4, 640, 10
    N = (H // 16) * (W // 16)
    torch.manual_seed(0)
    attn = FakeAttention(C, heads)
    x = torch.randn(2 * Fl, N, C)
    m32, m16 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)
    st = rp.StoryState(total_count=10 ** 9, height=H, width=W, mask1024=m32, mask4096=m16)
    st.write, st.cur_step = True, 25
    orc = rp.ConsistentAttnOracle(st, id_length=Fl)
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 10 ** 9
    host.mask1024, host.mask4096 = m32.to(DEV), m16.to(DEV)
    host.write, host.cur_step = True, 25
    proc = make_processor_class(host)(id_length=Fl)
    assert proc.native_projections
    gattn = FakeAttention(C, heads)
    gattn.load_state_dict(attn.state_dict())
    gattn = gattn.to(DEV, torch.bfloat16)
    with torch.no_grad():
        random.seed(0)
        want = orc(attn, x)
        random.seed(0)
        before = dict(native.LAUNCHES)
        got = proc(gattn, x.to(DEV, torch.bfloat16))
    assert _gemms(native.LAUNCHES) - _gemms(before) == 2
    # SDXL shapes run on the hand-written GEMM (q|k|v in one launch, then out), and the consistent write pass needs
    # no gather launch: the projection's epilogue fills K[S] / V[S]
    assert native.LAUNCHES["csa_gemm"] - before["csa_gemm"] == 2
    assert native.LAUNCHES["csa_gather_kv"] == before["csa_gather_kv"]
    err, cos = max_abs_cos(got, want)
    assert err <= MAX_ABS and cos >= MIN_COS, f"max-abs {err:.3e} cos {cos:.6f}"
