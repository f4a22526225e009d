"""Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
The reference classes are imported through oracle/ref_loader.py (gradio/diffusers/spaces/cog stubbed) and driven
with the stand-in ``Attention`` of oracle/fake_diffusers.py.  Outputs:

  masks.npz      cal_attn_mask_xl (StoryDiffusion/utils/gradio_utils.py:241-287) under fixed torch seeds: the
                 distinct rows of both masks (bit-packed), their index lists, and the proof that every row of a
                 block equals the block's first row.
  story.npz      the state machine of SpatialAttnProcessor2_0.__call__ (Comic_Generation.py:74-127) driven exactly
                 like SURVEY.md Appendix C: setup_seed(2047), three processors (two /32-class, one /16-class),
                 total_count=3, 8 write steps (batch 8) then 8 read steps (batch 2): inputs, weights, outputs of
                 every call, gate draws, branch trace.
  calls.npz      direct __call1__/__call2__ calls with a supplied dense mask (write and read geometry).
"""
from __future__ import annotations

import copy
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.fake_diffusers import FakeAttention  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def gen_masks(ref):
    cases = [  # (seed, T, F, sa32, sa64, h, w, dtype)
        (0, 5, 4, 0.5, 0.5, 128, 128, "float32"),
        (1, 5, 4, 0.5, 0.5, 256, 256, "float16"),
        (2, 5, 4, 0.0, 1.0, 128, 128, "float32"),
        (3, 4, 3, 0.3, 0.7, 192, 160, "float32"),
        (2047, 5, 4, 0.5, 0.5, 768, 768, "float32"),
        (7, 3, 2, 0.9, 0.1, 96, 96, "float16"),
    ]
    out = {"n_cases": np.array(len(cases))}
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    for ci, (seed, T, Fl, sa32, sa64, h, w, dt) in enumerate(cases):
        torch.manual_seed(seed)
        m32, m16 = gu.cal_attn_mask_xl(T, Fl, sa32, sa64, h, w, device="cpu", dtype=getattr(torch, dt))
        out[f"c{ci}_params"] = np.array([seed, T, Fl, h, w], dtype=np.int64)
        out[f"c{ci}_sa"] = np.array([sa32, sa64], dtype=np.float64)
        out[f"c{ci}_dtype"] = np.array(dt)
        for tag, m in (("32", m32), ("16", m16)):
            n = m.shape[0] // T
            rows = m[::n]  # the T distinct rows
            # premise of the compaction, verified on the real mask
            uniform = bool((m.view(T, n, -1) == rows.unsqueeze(1)).all())
            assert uniform
            out[f"c{ci}_rows{tag}"] = np.packbits(rows.numpy().astype(np.uint8), axis=1)
            out[f"c{ci}_n{tag}"] = np.array(n)
            out[f"c{ci}_counts{tag}"] = rows.sum(1).numpy().astype(np.int32)
            for f in range(T):
                out[f"c{ci}_idx{tag}_{f}"] = torch.nonzero(m[f * n], as_tuple=True)[0].numpy().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)
    print("masks.npz:", len(cases), "cases")


def gen_story(ref):
    H = W = 64
    Fl, C, heads = 4, 128, 2
    n32, n16 = (H // 32) * (W // 32), (H // 16) * (W // 16)
    layer_tokens = [n32, n32, n16]
    steps = 8
    ref.setup_seed(2047)
    attns = [FakeAttention(C, heads) for _ in layer_tokens]
    # inputs from a private generator so that the global torch stream is consumed only by the reference
    g = torch.Generator().manual_seed(1234)
    hs_w = [[torch.randn(2 * Fl, n, C, generator=g) for n in layer_tokens] for _ in range(steps)]
    hs_r = [[torch.randn(2, n, C, generator=g) for n in layer_tokens] for _ in range(steps)]

    procs = copy.deepcopy([ref.SpatialAttnProcessor2_0(id_length=Fl, device="cpu", dtype=torch.float32)
                           for _ in layer_tokens])
    ref.total_count, ref.attn_count, ref.cur_step = len(procs), 0, 0
    ref.sa32 = ref.sa64 = 0.5
    ref.height, ref.width = H, W
    ref.id_length, ref.total_length = Fl, Fl + 1
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    ref.mask1024, ref.mask4096 = gu.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device="cpu", dtype=torch.float32)

    draws = []
    orig_random = random.random

    def traced_random():
        v = orig_random()
        draws.append(v)
        return v

    trace = []
    for p in procs:
        c1, c2 = p.__call1__, p.__call2__
        p.__call1__ = (lambda f: lambda *a, **k: (trace.append(("consistent", ref.cur_step)), f(*a, **k))[1])(c1)
        p.__call2__ = (lambda f: lambda *a, **k: (trace.append(
            ("standard", ref.cur_step, a[2] is not None)), f(*a, **k))[1])(c2)
    random.random = traced_random
    out = {}
    try:
        with torch.no_grad():
            ref.write = True
            ref.cur_step = 0
            for s in range(steps):
                for li, p in enumerate(procs):
                    o = p(attns[li], hs_w[s][li])
                    out[f"w{s}_{li}_out"] = o.numpy().copy()
                    out[f"w{s}_{li}_in"] = hs_w[s][li].numpy().copy()
            bank_keys = [sorted(p.id_bank.keys()) for p in procs]
            ref.write = False
            ref.cur_step = 0
            for s in range(steps):
                for li, p in enumerate(procs):
                    o = p(attns[li], hs_r[s][li])
                    out[f"r{s}_{li}_out"] = o.numpy().copy()
                    out[f"r{s}_{li}_in"] = hs_r[s][li].numpy().copy()
    finally:
        random.random = orig_random
    for li, a in enumerate(attns):
        for k, v in a.state_dict().items():
            out[f"attn{li}_{k}"] = v.numpy().copy()
    out["draws"] = np.array(draws, dtype=np.float64)
    out["trace"] = np.array([f"{t[0]}:{t[1]}:{int(t[2]) if len(t) > 2 else -1}" for t in trace])
    out["bank_keys"] = np.array(bank_keys, dtype=np.int64)
    out["geom"] = np.array([H, W, Fl, C, heads, steps], dtype=np.int64)
    out["final_cur_step"] = np.array(ref.cur_step)
    np.savez_compressed(os.path.join(OUT, "story.npz"), **out)
    print("story.npz: draws", [round(d, 4) for d in draws])
    print("           trace", len(trace), "calls; bank keys", bank_keys[0])


def gen_calls(ref):
    """Direct __call1__ / __call2__ with a supplied dense mask: write (mask[:FN,:FN]) and read (mask[FN:])."""
    H = W = 128
    Fl, C, heads = 4, 128, 2
    n16 = (H // 16) * (W // 16)  # 64 tokens
    torch.manual_seed(11)
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    _, m16 = gu.cal_attn_mask_xl(Fl + 1, Fl, 0.5, 0.5, H, W, device="cpu", dtype=torch.float32)
    attn = FakeAttention(C, heads)
    proc = ref.SpatialAttnProcessor2_0(id_length=Fl, device="cpu", dtype=torch.float32)
    g = torch.Generator().manual_seed(99)
    hs_w = torch.randn(2 * Fl, n16, C, generator=g)
    hs_r = torch.randn(2, n16, C, generator=g)
    cut = m16.shape[0] // (Fl + 1) * Fl
    out = {}
    with torch.no_grad():
        out["write_consistent"] = proc.__call1__(attn, hs_w, None, m16[:cut, :cut]).numpy()
        out["write_standard"] = proc.__call2__(attn, hs_w, None, None).numpy()
        enc = torch.cat((hs_w[:Fl], hs_r[:1], hs_w[Fl:], hs_r[1:]))
        out["read_consistent"] = proc.__call1__(attn, hs_r, enc, m16[cut:]).numpy()
        out["read_early"] = proc.__call2__(attn, hs_r, enc, None).numpy()
    out["hs_w"], out["hs_r"] = hs_w.numpy(), hs_r.numpy()
    n = n16
    out["rows16"] = np.packbits(m16[::n].numpy().astype(np.uint8), axis=1)
    for k, v in attn.state_dict().items():
        out[f"attn_{k}"] = v.numpy().copy()
    out["geom"] = np.array([H, W, Fl, C, heads], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "calls.npz"), **out)
    print("calls.npz")


if __name__ == "__main__":
    torch.set_num_threads(4)
    ref = ref_loader.load_reference()
    gen_masks(ref)
    gen_story(ref)
    gen_calls(ref)
