"""Generate tests/golden/stock.npz by running the UNMODIFIED stock processor of the reference on CPU:
``AttnProcessor2_0`` of StoryDiffusion/utils/gradio_utils.py:387-472, imported through oracle/ref_loader.py and
driven with the stand-in ``Attention`` of oracle/fake_diffusers.py.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden_stock.py

Cases: self-attention (3-D and 4-D input, residual connection + rescale factor), cross-attention with a 77-token
encoder sequence of another width (to_k / to_v: cross_dim -> C), a ragged sequence length.  Recorded per case: the
module's weights, the inputs and the reference's output."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.fake_diffusers import FakeAttention  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, B, N (or (H, W) for 4-D), C, heads, cross (tokens, dim) or None, residual, rescale
    ("self3d", 2, 160, 128, 2, None, False, 1.0),
    ("self4d", 2, (12, 8), 128, 2, None, True, 2.0),
    ("cross", 3, 136, 128, 2, (77, 256), False, 1.0),
    ("cross_ragged", 2, 100, 64, 1, (77, 128), True, 1.0),
]


def make_attn(C, heads, cross, residual, rescale, gen):
    attn = FakeAttention(C, heads)
    if cross is not None:
        attn.to_k = torch.nn.Linear(cross[1], C, bias=False)
        attn.to_v = torch.nn.Linear(cross[1], C, bias=False)
    attn.residual_connection = residual
    attn.rescale_output_factor = rescale
    with torch.no_grad():
        for p in attn.parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * p.shape[-1] ** -0.5)
    return attn


def main():
    ref_loader.load_reference()
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    proc = gu.AttnProcessor2_0()
    out = {"n_cases": np.array(len(CASES))}
    for ci, (name, B, N, C, heads, cross, residual, rescale) in enumerate(CASES):
        gen = torch.Generator().manual_seed(100 + ci)
        attn = make_attn(C, heads, cross, residual, rescale, gen)
        if isinstance(N, tuple):
            x = torch.randn((B, C, N[0], N[1]), generator=gen)
        else:
            x = torch.randn((B, N, C), generator=gen)
        enc = None if cross is None else torch.randn((B, cross[0], cross[1]), generator=gen)
        with torch.no_grad():
            y = proc(attn, x.clone(), encoder_hidden_states=enc)
        out[f"c{ci}_name"] = np.array(name)
        out[f"c{ci}_x"] = x.numpy()
        if enc is not None:
            out[f"c{ci}_enc"] = enc.numpy()
        out[f"c{ci}_y"] = y.numpy()
        for k, p in attn.state_dict().items():
            out[f"c{ci}_w_{k}"] = p.numpy()
    np.savez_compressed(os.path.join(OUT, "stock.npz"), **out)
    print("stock.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
