#!/usr/bin/env python
"""Same-box GPU comparator: what library kernels make of the attention call on this
B200.  Two problems, per SDXL layer class of the bench workload:

* dense  — the REFERENCE's call as written: `F.scaled_dot_product_attention(q, k, v, attn_mask=dense_bool)` over all
  frames' tokens of a CFG half (StoryDiffusion/Comic_Generation.py:175-177), per torch SDPA backend;
* compact — the identical problem our kernel solves: per (CFG half, frame) the frame's N queries against its
  pre-gathered key list (K_f rows, gather outside the timed region), no mask — torch SDPA FLASH / CUDNN / EFFICIENT
  backends called per frame, and flash-attn 2.8 `flash_attn_varlen_func` as ONE launch over all 2F units.

Reported next to our kernel's numbers (`csa` entry: csa_attn_fwd on the same tensors); library kernels are not part
of the product path.

    python tools/bench_torch_sdpa.py [--frames 4] [--res 1024] [--sa 0.5] [--dtype bf16]
Prints one JSON line: per layer class and kernel ms per call and ALGORITHMIC TFLOP/s (the FLOP count bench.py uses:
only the keys the mask keeps), and the 30 + 6 layer step equivalent of the fastest library kernel of each problem."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spider_b200 import masks as csa_masks  # noqa: E402
from spider_b200 import native  # noqa: E402


def timeit(fn, iters, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--sa", type=float, default=0.5)
    ap.add_argument("--dtype", choices=["bf16", "fp16"], default="bf16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-mask problem (F >= 16: the mask is GBs)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype]
    Fl, T = args.frames, args.frames + 1
    from torch.nn.attention import SDPBackend, sdpa_kernel
    dense_backends = {"efficient": SDPBackend.EFFICIENT_ATTENTION, "cudnn": SDPBackend.CUDNN_ATTENTION,
                      "math": SDPBackend.MATH}
    compact_backends = {"torch_flash": SDPBackend.FLASH_ATTENTION, "torch_cudnn": SDPBackend.CUDNN_ATTENTION,
                        "torch_efficient": SDPBackend.EFFICIENT_ATTENTION}
    out = {"what": "library attention kernels on this GPU: the reference's dense-masked call, and the compact "
                   "per-frame problem csa_attn_fwd solves (keys pre-gathered outside the timed region)",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "frames": Fl, "res": args.res, "sa": args.sa, "dtype": args.dtype, "layers": {}}
    best = {"dense": {}, "compact": {}, "csa": {}}
    classes = [("32x32", (args.res // 32) ** 2, 1280, 20, 30), ("64x64", (args.res // 16) ** 2, 640, 10, 6)]
    for name, N, C, heads, count in classes:
        torch.manual_seed(0)
        sample = torch.rand((T * N,), device=dev) < args.sa
        cm = csa_masks.CompactMask(T, Fl, N, sample=sample)
        idx, counts = cm.lists(dev)
        kf_list = [int(c) for c in counts[:Fl].tolist()]
        kf = sum(kf_list)                                           # sum_f K_f (one list per frame)
        flops = 4 * 64 * heads * 2 * N * kf
        entry = {"tokens": N, "channels": C, "heads": heads, "per_step": count,
                 "algorithmic_tflop": round(flops / 1e12, 4), "dense": {}, "compact": {}}

        # ---------------------------------------------------------------- our kernel on the same tensors
        q2 = torch.randn(2 * Fl * N, C, device=dev, dtype=dtype)
        k2, v2 = torch.randn_like(q2), torch.randn_like(q2)
        o2 = torch.empty_like(q2)
        s_idx, s_count, ranges = cm.sample_list(dev)
        k_s, v_s, cap = native.gather_kv(k2, v2, Fl * N, 2, s_idx, s_count, Fl * N)
        ms = timeit(lambda: native.attn_fwd(q2, o2, heads=heads, n_groups=2, n_frames=Fl, n_q=N, k_a=k_s, v_a=v_s,
                                            a_group_rows=cap, ranges=ranges, range_base=0, range_step=1, k_b=k2,
                                            v_b=v2, b_group_rows=Fl * N, cb=(0, N, N)), args.iters, warm=3)
        entry["csa"] = {"ms": round(ms, 4), "tflops": round(flops / ms * 1e-9, 1)}
        best["csa"][name] = ms * count

        # ---------------------------------------------------------------- compact problem on library kernels
        # per unit (g, f): q (1, H, N, 64), k/v (1, H, K_f, 64), head-major contiguous as the libraries like it
        units = []
        for g in range(2):
            for f in range(Fl):
                keys = idx[f, :kf_list[f]].long() + g * Fl * N
                rows = slice((g * Fl + f) * N, (g * Fl + f + 1) * N)
                qh = q2[rows].view(N, heads, 64)
                kh = k2[keys].view(-1, heads, 64)
                vh = v2[keys].view(-1, heads, 64)
                units.append((qh, kh, vh))
        ref_unit = None
        for bname, b in compact_backends.items():
            try:
                bh = [(qh.transpose(0, 1)[None].contiguous(), kh.transpose(0, 1)[None].contiguous(),
                       vh.transpose(0, 1)[None].contiguous()) for qh, kh, vh in units]
                with sdpa_kernel(b):
                    def run():
                        for qq, kk, vv in bh:
                            F.scaled_dot_product_attention(qq, kk, vv)
                    ms = timeit(run, args.iters)
                    if ref_unit is None:
                        ref_unit = F.scaled_dot_product_attention(*bh[-1])[0].transpose(0, 1).reshape(N, C)
                entry["compact"][bname] = {"ms": round(ms, 4), "tflops": round(flops / ms * 1e-9, 1),
                                           "launches": len(bh)}
                del bh
            except Exception as e:   # noqa: BLE001  (a backend that does not take this call)
                entry["compact"][bname] = {"error": str(e).splitlines()[0][:160]}
                torch.cuda.synchronize()
        try:
            from flash_attn import flash_attn_varlen_func
            qv = torch.cat([u[0] for u in units])
            kv = torch.cat([u[1] for u in units])
            vv = torch.cat([u[2] for u in units])
            cu_q = torch.arange(0, (2 * Fl + 1) * N, N, device=dev, dtype=torch.int32)
            lens = torch.tensor([0] + kf_list * 2, device=dev, dtype=torch.int32)
            cu_k = torch.cumsum(lens, 0).to(torch.int32)
            ms = timeit(lambda: flash_attn_varlen_func(qv, kv, vv, cu_q, cu_k, N, max(kf_list)), args.iters)
            entry["compact"]["flash_attn_2.8_varlen"] = {"ms": round(ms, 4), "tflops": round(flops / ms * 1e-9, 1),
                                                         "launches": 1}
            fo = flash_attn_varlen_func(qv, kv, vv, cu_q, cu_k, N, max(kf_list))[-N:].reshape(N, C)
            if ref_unit is not None:
                entry["compact"]["flash_attn_2.8_varlen"]["max_abs_vs_torch"] = round(
                    (fo.float() - ref_unit.float()).abs().max().item(), 5)
            del qv, kv, vv
        except Exception as e:   # noqa: BLE001
            entry["compact"]["flash_attn_2.8_varlen"] = {"error": str(e).splitlines()[0][:160]}
            torch.cuda.synchronize()
        if ref_unit is not None:
            entry["csa"]["max_abs_vs_torch"] = round(
                (o2[-N:].float() - ref_unit.float()).abs().max().item(), 5)
        ok = [r["ms"] for r in entry["compact"].values() if "ms" in r]
        if ok:
            best["compact"][name] = min(ok) * count
        del units, q2, k2, v2, o2, k_s, v_s
        torch.cuda.empty_cache()

        # ---------------------------------------------------------------- the reference's dense-masked call
        if not args.no_dense:
            try:
                mask = cm.dense()[:Fl * N, :Fl * N].contiguous()    # mask[:F*N, :F*N] of the write pass (:105-114)
                q = torch.randn((2, heads, Fl * N, 64), device=dev, dtype=dtype)
                k, v = torch.randn_like(q), torch.randn_like(q)
                for bname, b in dense_backends.items():
                    try:
                        with sdpa_kernel(b):
                            ms = timeit(lambda: F.scaled_dot_product_attention(q, k, v, attn_mask=mask),
                                        max(2, args.iters // 2))
                        entry["dense"][bname] = {"ms": round(ms, 4), "tflops": round(flops / ms * 1e-9, 1)}
                    except Exception as e:   # noqa: BLE001
                        entry["dense"][bname] = {"error": str(e).splitlines()[0][:160]}
                        torch.cuda.synchronize()
                del q, k, v, mask
            except Exception as e:   # noqa: BLE001  (e.g. OOM building the (F*N)^2 mask)
                entry["dense"]["error"] = str(e).splitlines()[0][:160]
            ok = [r["ms"] for r in entry["dense"].values() if isinstance(r, dict) and "ms" in r]
            if ok:
                best["dense"][name] = min(ok) * count
        torch.cuda.empty_cache()
        out["layers"][name] = entry
    tflop = sum(out["layers"][n]["algorithmic_tflop"] * c for n, _, _, _, c in classes)
    out["step_equivalent"] = {}
    for prob, d in best.items():
        if len(d) == len(classes):
            total = sum(d.values())
            out["step_equivalent"][prob] = {"attention_ms": round(total, 3), "tflops": round(tflop / total * 1e3, 1)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
