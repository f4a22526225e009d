"""Shared helpers for the tests (fixture loading, FakeAttention reconstruction)."""
import os

import numpy as np
import torch

from oracle.fake_diffusers import FakeAttention

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def attn_from_fixture(z, prefix, channels, heads, dtype=torch.float32, device="cpu"):
    a = FakeAttention(channels, heads)
    sd = {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}
    a.load_state_dict(sd)
    return a.to(device=device, dtype=dtype)


def unpack_rows(packed, n_cols):
    return np.unpackbits(packed, axis=1)[:, :n_cols].astype(bool)


def max_abs_cos(a: torch.Tensor, b: torch.Tensor):
    a = a.detach().float().cpu().flatten()
    b = b.detach().float().cpu().flatten()
    err = (a - b).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    return err, cos


# tolerance of BASELINE.json's north_star for attention outputs vs the reference's fp32 output
MAX_ABS = 2e-2
MIN_COS = 0.9995
