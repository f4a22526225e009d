"""The hand-written sm_100a projection GEMM (csrc/gemm_sm100.cu, `-m gpu` on a B200) against a plain PyTorch fp32
reference of the same op: y = alpha * x w^T + bias with nn.Linear layouts, ragged M, strided operands, both 16-bit
types, every SDXL attn1 shape; and the fused gather of the sampled K/V rows against csa_gather_kv (bit-exact: the
same values, stored twice)."""
import pytest
import torch

from spider_b200 import masks as csa_masks
from spider_b200 import native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(x, w, bias, alpha):
    y = alpha * (x.float() @ w.float().t())
    if bias is not None:
        y = y + bias.float()
    return y


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,k,bias", [(8192, 1280, 1280, False), (8192, 2560, 1280, False), (4096, 1280, 1280, True),
                                        (32768, 640, 640, True), (32768, 1280, 640, False), (1000, 640, 640, True),
                                        (130, 128, 64, False), (1, 256, 128, True), (4608, 2560, 1280, False)])
def test_gemm_matches_torch_fp32(dtype, m, n, k, bias):
    assert native.gemm_supported(m, n, k)
    g = torch.Generator(device=DEV).manual_seed(m + n + k)
    x = torch.randn((m, k), device=DEV, generator=g).to(dtype)
    w = (torch.randn((n, k), device=DEV, generator=g) * k ** -0.5).to(dtype)
    b = torch.randn((n,), device=DEV, generator=g).to(dtype) if bias else None
    y = native.gemm(x, w, b)
    torch.cuda.synchronize()
    want = _ref(x, w, b, 1.0)
    err = (y.float() - want).abs().max().item()
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3     # output rounding of a 16-bit y with |y| up to ~5
    assert err <= tol, f"max-abs {err:.3e}"
    # against the cuBLASLt path: same inputs, same fp32 accumulation, same 16-bit rounding of the result
    y2 = native.linear(x, w, b)
    torch.cuda.synchronize()
    assert (y.float() - y2.float()).abs().max().item() <= tol


def test_gemm_alpha_strides_and_out():
    dtype = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(1)
    big = torch.randn((700, 1280 + 64), device=DEV, generator=g).to(dtype)
    x = big[:, 64:]                                        # row stride 1344, 128-byte aligned start
    w = (torch.randn((1280, 1280), device=DEV, generator=g) * 0.03).to(dtype)
    out = torch.zeros((700, 2560), device=DEV, dtype=dtype)
    y = native.gemm(x, w, out=out[:, 1280:], alpha=0.125)
    torch.cuda.synchronize()
    assert y.data_ptr() == out[:, 1280:].data_ptr() and float(out[:, :1280].abs().max()) == 0.0
    want = _ref(x, w, None, 0.125)
    assert (y.float() - want).abs().max().item() <= 5e-3


@pytest.mark.parametrize("m,C", [(8192, 1280), (4096, 640), (1000, 640)])
def test_gemm_two_outputs_and_ragged_wide_tiles(m, C):
    """Stacked weight [w_q; w_k; w_v] (N = 3C; 1920 is not a multiple of the 256-wide tile): q lands in `out`, K|V in
    `out2`, both equal the separate products bit for bit."""
    dtype = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(m + C)
    x = torch.randn((m, C), device=DEV, generator=g).to(dtype)
    w = (torch.randn((3 * C, C), device=DEV, generator=g) * C ** -0.5).to(dtype)
    q = torch.full((m, C), 3.0, device=DEV, dtype=dtype)
    kv = torch.full((m, 2 * C), 3.0, device=DEV, dtype=dtype)
    native.gemm(x, w, out=q, out2=kv)
    q_ref = native.gemm(x, w[:C])
    kv_ref = native.gemm(x, w[C:])
    torch.cuda.synchronize()
    assert torch.equal(q, q_ref) and torch.equal(kv, kv_ref)
    want = _ref(x, w, None, 1.0)
    assert (torch.cat([q, kv], dim=1).float() - want).abs().max().item() <= 2e-2


def test_gemm_rejects_bad_arguments():
    x = torch.zeros((128, 64), device=DEV, dtype=torch.bfloat16)
    w = torch.zeros((128, 64), device=DEV, dtype=torch.bfloat16)
    assert not native.gemm_supported(128, 100, 64) and not native.gemm_supported(128, 128, 48)
    with pytest.raises(native.CsaNativeError):
        native.gemm(x, torch.zeros((100, 64), device=DEV, dtype=torch.bfloat16))
    with pytest.raises(native.CsaNativeError):
        native.gemm(x, w.to(torch.float16))
    native.abort_batch()


@pytest.mark.parametrize("Fl,N,C", [(4, 1024, 1280), (4, 4096, 640), (3, 576, 1280)])
def test_fused_gather_equals_gather_kv(Fl, N, C):
    """K|V projection with the fused gather: K[S] / V[S] filled by the epilogue are bit-identical to csa_gather_kv run
    on the projection's own output, for both CFG halves; rows that are not sampled are not touched."""
    dtype = torch.bfloat16
    T = Fl + 1
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn((2 * Fl * N, C), device=DEV, generator=g).to(dtype)
    w_kv = (torch.randn((2 * C, C), device=DEV, generator=g) * C ** -0.5).to(dtype)
    sample = torch.rand((T * N,), device=DEV, generator=g) < 0.5
    cm = csa_masks.CompactMask(T, Fl, N, sample=sample)
    s_idx, s_count, _ = cm.sample_list(DEV)
    pos = cm.sample_positions(DEV)
    cnt = int(s_count.item())
    # the inverse list: pos[s_idx[i]] == i, -1 elsewhere
    want_pos = torch.full((Fl * N,), -1, dtype=torch.int32, device=DEV)
    want_pos[s_idx[:cnt].long()] = torch.arange(cnt, dtype=torch.int32, device=DEV)
    assert torch.equal(pos, want_pos)
    cap = Fl * N + native.CSA_TILE
    k_s = torch.full((2 * cap, C), 7.0, device=DEV, dtype=dtype)
    v_s = torch.full((2 * cap, C), 7.0, device=DEV, dtype=dtype)
    kv = native.gemm(x, w_kv, scatter=(pos, k_s, v_s, Fl * N, cap, C))
    k_ref, v_ref, cap2 = native.gather_kv(kv[:, :C], kv[:, C:], Fl * N, 2, s_idx, s_count, Fl * N)
    # the same through the stacked q|k|v weight with two outputs (what the processor issues)
    w_q = (torch.randn((C, C), device=DEV, generator=g) * C ** -0.5).to(dtype)
    k_s2 = torch.full((2 * cap, C), 7.0, device=DEV, dtype=dtype)
    v_s2 = torch.full((2 * cap, C), 7.0, device=DEV, dtype=dtype)
    q3 = torch.empty((2 * Fl * N, C), device=DEV, dtype=dtype)
    kv3 = torch.empty((2 * Fl * N, 2 * C), device=DEV, dtype=dtype)
    native.gemm(x, torch.cat([w_q, w_kv]).contiguous(), out=q3, out2=kv3,
                scatter=(pos, k_s2, v_s2, Fl * N, cap, 2 * C, C))
    torch.cuda.synchronize()
    assert cap2 == cap
    assert torch.equal(kv3, kv) and torch.equal(k_s2, k_s) and torch.equal(v_s2, v_s)
    for grp in range(2):
        assert torch.equal(k_s[grp * cap:grp * cap + cnt], k_ref[grp * cap:grp * cap + cnt])
        assert torch.equal(v_s[grp * cap:grp * cap + cnt], v_ref[grp * cap:grp * cap + cnt])
        assert float((k_s[grp * cap + cnt:(grp + 1) * cap].float() - 7.0).abs().max()) == 0.0   # untouched


@pytest.mark.parametrize("n_peers", [1, 2])
def test_fused_exchange_on_one_gpu(n_peers):
    """csa_gemm(exchange=...) and csa_attn_fwd(done_dst=...) with every "peer" buffer on this GPU (the kernels only
    see pointers): the epilogue stores this rank's sampled rows into all destination buffers at their position in S
    and raises its arrival flag in each, over several epochs with device-resident epoch arithmetic; the attention
    launch waits for the flag, reproduces the single-GPU launch, and publishes the release flags itself."""
    dtype = torch.bfloat16
    Fl, N, C, heads = 4, 256, 128, 2
    T = Fl + 1
    me = n_peers - 1                       # this "rank" holds the last frames of the half
    fr = Fl // n_peers
    f0 = me * fr
    g = torch.Generator(device=DEV).manual_seed(11)
    sample = torch.rand((T * N,), device=DEV, generator=g) < 0.5
    cm = csa_masks.CompactMask(T, Fl, N, sample=sample)
    s_idx, s_count, ranges = cm.sample_list(DEV)
    pos = cm.sample_positions(DEV)
    rows = Fl * N + native.CSA_TILE
    w_qkv = (torch.randn((3 * C, C), device=DEV, generator=g) * C ** -0.5).to(dtype)
    flags = torch.zeros((n_peers, 3, native.CSA_MAX_PEERS), dtype=torch.int32, device=DEV)
    ready = [flags[r, 0] for r in range(n_peers)]
    done = [flags[r, 1] for r in range(n_peers)]
    counter, counter2 = flags[me, 2], flags[me, 2][1:]
    epoch_base = torch.full((1,), 40, dtype=torch.int32, device=DEV)
    for epoch in (1, 2, 3):
        x_full = torch.randn((Fl * N, C), device=DEV, generator=g).to(dtype)      # one CFG half, all frames
        # single-GPU result: fused local gather + one attention launch
        q1 = torch.empty((Fl * N, C), device=DEV, dtype=dtype)
        kv1 = torch.empty((Fl * N, 2 * C), device=DEV, dtype=dtype)
        k_s, v_s = (torch.zeros((rows, C), device=DEV, dtype=dtype) for _ in range(2))
        native.gemm(x_full, w_qkv, out=q1, out2=kv1, scatter=(pos, k_s, v_s, Fl * N, rows, 2 * C, C))
        want = native.attn_fwd(q1, torch.empty_like(q1), heads=heads, n_groups=1, n_frames=Fl, n_q=N, k_a=k_s,
                               v_a=v_s, a_group_rows=rows, ranges=ranges, range_base=0, range_step=1,
                               k_b=kv1[:, :C], v_b=kv1[:, C:], b_group_rows=Fl * N, cb=(0, N, N))
        # "sharded": this rank projects its own frames and delivers their sampled rows to every destination; the
        # other frames' rows are put there by hand (what the other ranks' launches would do) with their flags
        ks = [torch.full((rows, C), 7.0, device=DEV, dtype=dtype) for _ in range(n_peers)]   # finite: masked keys
        vs = [torch.full((rows, C), 7.0, device=DEV, dtype=dtype) for _ in range(n_peers)]   # still enter P V as 0 * v
        lo = int(ranges[f0][1].item()) if f0 > 0 else 0       # S position of this rank's first sampled row
        if f0 > 0:
            for r in range(n_peers):
                ks[r][:lo] = k_s[:lo]
                vs[r][:lo] = v_s[:lo]
                flags[r, 0, :me] = 40 + epoch
        x = x_full[f0 * N:].contiguous()
        q2 = torch.empty((fr * N, C), device=DEV, dtype=dtype)
        kv2 = torch.empty((fr * N, 2 * C), device=DEV, dtype=dtype)
        for r in range(n_peers):                               # the peers released the slot two epochs ago
            if r != me:
                done[me][r] = 40 + epoch - 2
        torch.cuda.synchronize()
        native.gemm(x, w_qkv, out=q2, out2=kv2, scatter=(pos[f0 * N:], None, None, fr * N, 0, 2 * C, C),
                    exchange={"k_dst": ks, "v_dst": vs, "ready": ready, "self": me, "epoch": epoch, "done": done[me],
                              "done_epoch": epoch - 2, "counter": counter, "epoch_base": epoch_base})
        o = native.attn_fwd(q2, torch.empty_like(q2), heads=heads, n_groups=1, n_frames=fr, n_q=N, k_a=ks[me],
                            v_a=vs[me], a_group_rows=rows, ranges=ranges, range_base=f0, range_step=1,
                            k_b=kv2[:, :C], v_b=kv2[:, C:], b_group_rows=fr * N, cb=(0, N, N), b_first=True,
                            ready=ready[me], ready_epoch=epoch, ready_peers=n_peers, ready_frames_per_peer=fr,
                            epoch_base=epoch_base, done=done, done_counter=counter2, peer_self=me)
        torch.cuda.synchronize()
        cnt = int(s_count.item())
        for r in range(n_peers):
            assert torch.equal(ks[r][lo:cnt], k_s[lo:cnt]) and torch.equal(vs[r][lo:cnt], v_s[lo:cnt])
            assert int(flags[r, 0, me]) == 40 + epoch                     # arrival flag raised everywhere
        assert int(counter[0]) == 0
        # same math; the key order (own block first) and the work decomposition differ: output rounding only
        assert (o.float() - want[f0 * N:].float()).abs().max().item() <= 4e-3
        for r in range(n_peers):
            if r != me:
                assert int(flags[r, 1, me]) == 40 + epoch                 # released by the attention launch
        assert int(counter2[0]) == 0
    assert native.debug_stuck() is None
