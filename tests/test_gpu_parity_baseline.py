"""GPU parity at BASELINE.json's own shapes, through the PROCESSOR (write, then read via the id_bank) against the CPU
oracle in fp32 on identical inputs, weights, masks and RNG state (run with `-m gpu` on a B200):

  * config 1 shape: the 64x64 SDXL layer at 1024^2 (N=4096, C=640, 10 heads), 4 frames x2 CFG, sa64 = 0.5;
  * 768^2 — the reference's default resolution (Comic_Generation.py:346-347): N=576 (C=1280, 20 heads) and N=2304
    (C=640, 10 heads), both ragged against the kernel's 128-row tiles;
  * config 3 placement: all 70 attn1 layers of an SDXL-shaped UNet swapped (install.set_attention_processor with
    all_self_attn), two denoise steps of the write pass and two of the read pass through the state machine;
  * config 4 shape: 16 frames, where the reference cannot build its dense mask — every (frame) unit of one CFG half
    against the oracle's gathered form (rp.gathered_attention).
Tolerance: max-abs <= 2e-2 and cosine >= 0.9995 against the fp32 oracle (BASELINE.json north_star).
"""
import copy
import random

import pytest
import torch

import spider_b200
from spider_b200 import masks as csa_masks
from spider_b200 import native
from spider_b200.install import make_processor_class, set_attention_processor
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention, FakeUNet, sdxl_layout

from helpers import MAX_ABS, MIN_COS, max_abs_cos

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _assert_close(got, want, what):
    err, cos = max_abs_cos(got, want)
    assert err <= MAX_ABS and cos >= MIN_COS, f"{what}: max-abs {err:.3e} cos {cos:.6f}"
    return err, cos


def _write_then_read(H, W, N, C, heads, Fl, use32, dtype, seed=0):
    """One layer: write-consistent call, then read-consistent call through the bank; product (GPU) vs oracle (CPU)."""
    torch.manual_seed(seed)
    random.seed(seed)
    attn = FakeAttention(C, heads)
    hs_w = torch.randn(2 * Fl, N, C)
    hs_r = torch.randn(2, N, C)
    m32, m16 = rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)
    mask = m32 if use32 else m16
    del m32, m16
    st = rp.StoryState(total_count=10 ** 9, height=H, width=W)
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.total_count = H, W, 10 ** 9
    if use32:
        st.mask1024, host.mask1024 = mask, mask.to(DEV)
    else:
        st.mask4096, host.mask4096 = mask, mask.to(DEV)
    orc = rp.ConsistentAttnOracle(st, id_length=Fl)
    with torch.no_grad():
        st.write, st.cur_step = True, 25
        random.seed(seed)
        want_w = orc(attn, hs_w)
        st.write = False
        random.seed(seed)
        want_r = orc(attn, hs_r)
    assert [t[0] for t in st.trace] == ["consistent", "consistent"]
    p = make_processor_class(host)(id_length=Fl)
    ga = copy.deepcopy(attn).to(DEV, dtype)
    with torch.no_grad():
        host.write, host.cur_step = True, 25
        random.seed(seed)
        got_w = p(ga, hs_w.to(DEV, dtype))
        assert p._last_branch == "consistent"
        host.write = False
        random.seed(seed)
        got_r = p(ga, hs_r.to(DEV, dtype))
        assert p._last_branch == "consistent"
    torch.cuda.synchronize()
    assert native.debug_stuck() is None
    return (got_w, want_w), (got_r, want_r)


def test_processor_vs_oracle_config1_shape():
    """BASELINE config 1: one SDXL 64x64 self-attention layer (640 ch, 10 heads), 4-frame story x2 CFG, sa64 = 0.5 —
    the oracle runs it in fp32 on the CPU exactly as the config states, the product in bf16 on the B200."""
    (gw, ww), (gr, wr) = _write_then_read(1024, 1024, 4096, 640, 10, 4, use32=False, dtype=torch.bfloat16)
    e, c = _assert_close(gw, ww, "config1 write 64x64")
    print(f"config1 write: max-abs {e:.3e} cos {c:.6f}")
    e, c = _assert_close(gr, wr, "config1 read 64x64")
    print(f"config1 read: max-abs {e:.3e} cos {c:.6f}")


@pytest.mark.parametrize("N,C,heads,use32,dtype", [(576, 1280, 20, True, torch.bfloat16),
                                                    (2304, 640, 10, False, torch.bfloat16),
                                                    (576, 1280, 20, True, torch.float16),
                                                    (2304, 640, 10, False, torch.float16)])
def test_processor_vs_oracle_768(N, C, heads, use32, dtype):
    """768x768 (the reference's default, Comic_Generation.py:346-347): one layer of each class, write then read."""
    (gw, ww), (gr, wr) = _write_then_read(768, 768, N, C, heads, 4, use32=use32, dtype=dtype, seed=1)
    _assert_close(gw, ww, f"768^2 write N={N}")
    _assert_close(gr, wr, f"768^2 read N={N}")


def test_all_70_placement_two_steps_vs_oracle():
    """BASELINE config 3 placement: every attn1 layer of an SDXL-shaped UNet swapped (70 processors, total_count 70),
    two denoise steps of the write pass then two of the read pass through the whole state machine (gate draws, step
    roll-over, mask re-sampling from the torch generator) — every call's output against the oracle's, same branch
    trace, same final state.  256^2 latents so that the CPU oracle finishes in seconds."""
    H = W = 256
    Fl = 4
    n32, n16 = (H // 32) * (W // 32), (H // 16) * (W // 16)
    torch.manual_seed(11)
    unet = FakeUNet(sdxl_layout())
    layers = unet.self_attn_layers()
    assert len(layers) == 70
    inputs_w = [torch.randn(2 * Fl, n32 if a.inner_dim == 1280 else n16, a.inner_dim) for _, a in layers]
    inputs_r = [torch.randn(2, n32 if a.inner_dim == 1280 else n16, a.inner_dim) for _, a in layers]

    def drive(call, state, sample_masks):
        """two write steps + two read steps starting at cur_step 24; returns outputs and the gate draws"""
        outs, draws = [], []
        orig = random.random
        random.random = lambda: (draws.append(orig()), draws[-1])[1]
        try:
            rp.setup_seed(99)
            state.mask1024, state.mask4096 = sample_masks()
            with torch.no_grad():
                for write, xs in ((True, inputs_w), (False, inputs_r)):
                    state.write, state.cur_step, state.attn_count = write, 24, 0
                    for _ in range(2):
                        for li, x in enumerate(xs):
                            outs.append(call(li, x))
        finally:
            random.random = orig
        return outs, draws

    # oracle: one processor per layer, like the reference's installation (:353-371)
    st = rp.StoryState(total_count=70, height=H, width=W)
    orcs = [rp.ConsistentAttnOracle(st, id_length=Fl) for _ in layers]
    want, draws_o = drive(lambda li, x: orcs[li](layers[li][1], x), st,
                          lambda: rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W))
    # product: processors installed by the product's own helper on a GPU copy of the UNet; they sample their masks on
    # the CPU generator in fp32 like the oracle does, so that both walk through the same sample vectors
    host = spider_b200.StoryGlobals()
    host.height, host.width, host.sa32, host.sa64 = H, W, 0.5, 0.5
    gunet = copy.deepcopy(unet).to(DEV, torch.bfloat16)
    cls = make_processor_class(host)
    n = set_attention_processor(gunet, Fl, host=host, all_self_attn=True, processor_cls=cls)
    assert n == 70 and host.total_count == 70
    glayers = gunet.self_attn_layers()
    for _, a in glayers:
        a.processor.device, a.processor.dtype = "cpu", torch.float32
    got, draws_p = drive(lambda li, x: glayers[li][1](x.to(DEV, torch.bfloat16)), host,
                         lambda: tuple(m.to(DEV) for m in rp.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W)))
    torch.cuda.synchronize()
    assert draws_p == draws_o
    assert host.cur_step == st.cur_step == 26 and host.attn_count == st.attn_count == 0
    worst = (0.0, 1.0)
    for i, (g, w_) in enumerate(zip(got, want)):
        e, c = _assert_close(g, w_, f"all-70 call {i}")
        worst = (max(worst[0], e), min(worst[1], c))
    branches = [t[0] for t in st.trace]
    assert branches.count("consistent") > 200      # 280 calls, gate 0.1 at cur_step >= 20
    print(f"all-70 placement, 280 calls ({branches.count('consistent')} consistent): worst max-abs {worst[0]:.3e} "
          f"cos {worst[1]:.6f}")


def test_f16_all_units_of_one_cfg_half_vs_oracle_gathered():
    """BASELINE config 4 shape: a 16-frame write pass of the 64x64 layer (the reference's dense mask would be 4.85 GB):
    the kernel's output for EVERY frame of one CFG half against the oracle's gathered form in fp32 on the CPU."""
    Fl, N, C, heads = 16, 4096, 640, 10
    T = Fl + 1
    torch.manual_seed(4)
    sample = torch.rand((T * N,)) < 0.5
    rows = rp.frame_rows(sample.clone(), T, Fl)[:Fl, :Fl * N]
    key_lists = rp.index_lists(rows)
    g = 1                                                            # the cond half
    q = torch.randn(2 * Fl * N, C).to(torch.bfloat16)
    k = torch.randn(2 * Fl * N, C).to(torch.bfloat16)
    v = torch.randn(2 * Fl * N, C).to(torch.bfloat16)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    cm = csa_masks.CompactMask(T, Fl, N, sample=sample.to(DEV))
    idx, counts = cm.lists(DEV)
    for f in range(Fl):                                               # index lists bit-exact
        c = int(counts[f])
        assert c == key_lists[f].numel() and torch.equal(idx[f, :c].cpu(), key_lists[f])
    s_idx, s_count, ranges = cm.sample_list(DEV)
    k_s, v_s, cap = native.gather_kv(kd, vd, Fl * N, 2, s_idx, s_count, Fl * N)
    o = torch.empty_like(qd)
    native.attn_fwd(qd, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k_s, v_a=v_s, a_group_rows=cap,
                    ranges=ranges, range_base=0, range_step=1, k_b=kd, v_b=vd, b_group_rows=Fl * N, cb=(0, N, N))
    torch.cuda.synchronize()
    assert native.debug_stuck() is None
    half = slice(g * Fl * N, (g + 1) * Fl * N)
    want = rp.gathered_attention(q[half][None].float(), k[half][None].float(), v[half][None].float(), key_lists,
                                 heads)[0]
    got = o[half].float().cpu()
    worst = (0.0, 1.0)
    for f in range(Fl):
        rs = slice(f * N, (f + 1) * N)
        e, c = _assert_close(got[rs], want[rs], f"F=16 frame {f}")
        worst = (max(worst[0], e), min(worst[1], c))
    print(f"F=16 all 16 units of the cond half: worst max-abs {worst[0]:.3e} cos {worst[1]:.6f}")


def test_bank_arena_capacity_and_equivalence():
    """bank_capacity=N: the K|V GEMM of a write step lands in a preallocated arena slot (same results as the growing
    bank, bit for bit), a step beyond the capacity raises instead of allocating, clear_bank() frees the slots."""
    from spider_b200.processor import BankCapacityError

    H = W = 256
    Fl, N, C, heads = 4, 64, 1280, 20
    torch.manual_seed(2)
    attn = FakeAttention(C, heads).to(DEV, torch.bfloat16)
    xs = [torch.randn(2 * Fl, N, C, device=DEV, dtype=torch.bfloat16) for _ in range(3)]
    xr = torch.randn(2, N, C, device=DEV, dtype=torch.bfloat16)
    outs = {}
    real = random.random
    random.random = lambda: 0.99
    try:
        for cap in (None, 2):
            host = spider_b200.StoryGlobals()
            host.height, host.width, host.total_count = H, W, 10 ** 9
            torch.cuda.manual_seed(9)
            host.mask1024, host.mask4096 = spider_b200.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device=DEV)
            cls = make_processor_class(host)
            cls.bank_capacity = cap
            p = cls(id_length=Fl, device=DEV)
            with torch.no_grad():
                host.write = True
                res = []
                for s in range(2):
                    host.cur_step = 25 + s
                    res.append(p(attn, xs[s]))
                if cap is not None:
                    host.cur_step = 27
                    with pytest.raises(BankCapacityError):
                        p(attn, xs[2])
                    native.abort_batch()
                    assert p.id_bank[25].k.data_ptr() == p._arena[0][0].data_ptr()
                host.write = False
                for s in range(2):
                    host.cur_step = 25 + s
                    res.append(p(attn, xr))
                if cap is not None:
                    p.clear_bank()
                    host.write, host.cur_step = True, 27
                    p(attn, xs[2])                       # fits again
            torch.cuda.synchronize()
            outs[cap] = res
    finally:
        random.random = real
    for a, b in zip(outs[None], outs[2]):
        assert torch.equal(a, b)
