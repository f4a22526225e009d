"""Quick on-GPU check of libcsa_b200.so against plain torch (debug tool; the real parity tests are tests/).

Runs compaction and the four attention modes on small and BASELINE shapes, prints max-abs error / cosine vs an
fp32 gathered-SDPA reference and CUDA-event timings.  Usage: python tools/gpu_check.py [--quick]
"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import native  # noqa: E402

dev = torch.device("cuda:0")


def sample_lists(F, N, sa, T=None):
    """reference mask algebra (gradio_utils.py:257-286) as index lists; rows 0..F-1 write, row F read"""
    T = T or F + 1
    r = (torch.rand((1, T * N), device=dev, dtype=torch.float16) < sa)[0]
    rows = []
    for i in range(T):
        row = r.clone()
        row[F * N:] = False
        row[i * N:(i + 1) * N] = True
        rows.append(row)
    return r, rows


def ref_attention(q, k, v, heads):
    # q (nq, C), k/v (nk, C) -> (nq, C); fp32
    nq, C = q.shape
    d = C // heads
    qh = q.float().view(nq, heads, d).transpose(0, 1)
    kh = k.float().view(-1, heads, d).transpose(0, 1)
    vh = v.float().view(-1, heads, d).transpose(0, 1)
    o = torch.nn.functional.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0]
    return o.transpose(0, 1).reshape(nq, C)


def report(name, out, ref):
    err = (out.float() - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
    ok = err <= 2e-2 and cos >= 0.9995
    print(f"  {name:<46s} max-abs {err:.3e}  cos {cos:.6f}  {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check_compact():
    print("compaction:")
    ok = True
    for (F, N, sa) in [(4, 64, 0.5), (4, 1024, 0.5), (4, 4096, 0.5), (3, 576, 0.3), (4, 100, 0.0), (4, 100, 1.0),
                       (16, 4096, 0.5)]:
        T = F + 1
        r, rows = sample_lists(F, N, sa)
        idx, counts = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=F * N)
        torch.cuda.synchronize()
        good = True
        for i in range(T):
            ref = torch.nonzero(rows[i])[:, 0].int()
            c = int(counts[i])
            good &= (c == ref.numel()) and torch.equal(idx[i, :c], ref)
        # dense-row path
        dense = torch.stack(rows)  # (T, T*N) bool
        idx2, counts2 = native.compact_rows(dense, T, T * N, dense.stride(0))
        torch.cuda.synchronize()
        for i in range(T):
            c = int(counts2[i])
            good &= (c == int(counts[i])) and torch.equal(idx2[i, :c], idx[i, :c])
        print(f"  F={F} N={N} sa={sa}: {'OK' if good else 'FAIL'} counts={counts.tolist()[:6]}")
        ok &= good
    return ok


def run_write(F, N, C, heads, sa, dtype, timing=False):
    G = 2
    q = torch.randn(G * F * N, C, device=dev, dtype=dtype)
    k = torch.randn(G * F * N, C, device=dev, dtype=dtype)
    v = torch.randn(G * F * N, C, device=dev, dtype=dtype)
    o = torch.empty_like(q)
    r, rows = sample_lists(F, N, sa)
    T = F + 1
    idx, counts = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=F * N)
    fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=G, n_frames=F, n_q=N, k_a=k, v_a=v,
                                 a_group_rows=F * N, idx=idx, counts=counts, list_base=0, list_step=1)
    fn()
    torch.cuda.synchronize()
    ok = True
    for g in range(G):
        for f in (0, F - 1):
            keys = torch.nonzero(rows[f][:F * N])[:, 0] + g * F * N
            qs = slice((g * F + f) * N, (g * F + f + 1) * N)
            ref = ref_attention(q[qs], k[keys], v[keys], heads)
            ok &= report(f"write F={F} N={N} C={C} sa={sa} {str(dtype)[6:]} g{g} f{f}", o[qs], ref)
    if timing:
        ms = time_it(fn)
        kf = counts[:F].sum().item()
        flops = 4 * 64 * heads * G * N * kf
        print(f"  -> {ms:.3f} ms  {flops / ms * 1e-9:.1f} TFLOP/s (algorithmic)")
    return ok


def run_standard(B, N, C, heads, dtype, timing=False):
    q = torch.randn(B * N, C, device=dev, dtype=dtype)
    k = torch.randn(B * N, C, device=dev, dtype=dtype)
    v = torch.randn(B * N, C, device=dev, dtype=dtype)
    o = torch.empty_like(q)
    fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=B, n_q=N, k_b=k, v_b=v,
                                 b_group_rows=B * N, cb=(0, N, N))
    fn()
    torch.cuda.synchronize()
    ok = True
    for b in (0, B - 1):
        s = slice(b * N, (b + 1) * N)
        ok &= report(f"standard B={B} N={N} C={C} {str(dtype)[6:]} b{b}", o[s], ref_attention(q[s], k[s], v[s], heads))
    if timing:
        ms = time_it(fn)
        flops = 4 * 64 * heads * B * N * N
        print(f"  -> {ms:.3f} ms  {flops / ms * 1e-9:.1f} TFLOP/s")
    return ok


def run_read(F, N, C, heads, sa, dtype, early=False, timing=False):
    G = 2
    q = torch.randn(G * N, C, device=dev, dtype=dtype)
    kc = torch.randn(G * N, C, device=dev, dtype=dtype)
    vc = torch.randn(G * N, C, device=dev, dtype=dtype)
    kb = torch.randn(G * F * N, C, device=dev, dtype=dtype)
    vb = torch.randn(G * F * N, C, device=dev, dtype=dtype)
    o = torch.empty_like(q)
    r, rows = sample_lists(F, N, sa)
    T = F + 1
    idx, counts = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=F * N)
    if early:
        fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=G, n_frames=1, n_q=N, k_a=kb, v_a=vb,
                                     a_group_rows=F * N, ca=(0, 0, F * N), k_b=kc, v_b=vc, b_group_rows=N,
                                     cb=(0, 0, N))
    else:
        fn = lambda: native.attn_fwd(q, o, heads=heads, n_groups=G, n_frames=1, n_q=N, k_a=kb, v_a=vb,
                                     a_group_rows=F * N, idx=idx, counts=counts, list_base=F, list_step=0,
                                     g_adjust=-N, k_b=kc, v_b=vc, b_group_rows=N, cb=(0, 0, N))
    fn()
    torch.cuda.synchronize()
    ok = True
    for g in range(G):
        if early:
            keys_b = torch.arange(F * N, device=dev) + g * F * N
        else:
            keys_b = torch.nonzero(rows[F][:F * N])[:, 0] + g * F * N
        kk = torch.cat([kb[keys_b], kc[g * N:(g + 1) * N]])
        vv = torch.cat([vb[keys_b], vc[g * N:(g + 1) * N]])
        s = slice(g * N, (g + 1) * N)
        ok &= report(f"read{'-early' if early else ''} F={F} N={N} C={C} sa={sa} {str(dtype)[6:]} g{g}", o[s],
                     ref_attention(q[s], kk, vv, heads))
    if timing:
        ms = time_it(fn)
        print(f"  -> {ms:.3f} ms")
    return ok


def main():
    quick = "--quick" in sys.argv
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    ok = check_compact()
    print("attention (small):", flush=True)
    ok &= run_standard(2, 256, 64, 1, torch.bfloat16)
    ok &= run_standard(3, 384, 128, 2, torch.float16)
    ok &= run_write(2, 256, 64, 1, 0.5, torch.bfloat16)
    ok &= run_write(4, 256, 640, 10, 0.5, torch.bfloat16)
    ok &= run_write(3, 576, 128, 2, 0.3, torch.float16)
    ok &= run_write(4, 100, 64, 1, 0.5, torch.bfloat16)
    ok &= run_write(4, 256, 128, 2, 0.0, torch.bfloat16)
    ok &= run_write(4, 256, 128, 2, 1.0, torch.bfloat16)
    ok &= run_read(4, 256, 128, 2, 0.5, torch.bfloat16)
    ok &= run_read(4, 576, 128, 2, 0.5, torch.float16)
    ok &= run_read(4, 256, 128, 2, 0.5, torch.bfloat16, early=True)
    if not quick:
        print("attention (BASELINE shapes):", flush=True)
        ok &= run_write(4, 4096, 640, 10, 0.5, torch.bfloat16, timing=True)
        ok &= run_write(4, 1024, 1280, 20, 0.5, torch.bfloat16, timing=True)
        ok &= run_write(4, 4096, 640, 10, 0.5, torch.float16, timing=True)
        ok &= run_standard(8, 4096, 640, 10, torch.bfloat16, timing=True)
        ok &= run_read(4, 4096, 640, 10, 0.5, torch.bfloat16, timing=True)
        ok &= run_read(4, 1024, 1280, 20, 0.5, torch.bfloat16, timing=True)
        ok &= run_write(4, 4096, 640, 10, 1.0, torch.bfloat16, timing=True)
    print("ALL OK" if ok else "SOME FAILED")
    st = native.debug_stuck()
    if st:
        print("watchdog record:", [hex(x) for x in st])
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
