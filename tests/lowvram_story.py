"""Replays the scenario of tests/golden/lowvram.npz (see tests/golden/make_golden_lowvram.py) through any processor
class with the low-VRAM control surface; shared by the CPU (oracle) and GPU (B200 processor) tests."""
import random

import numpy as np
import torch

from helpers import attn_from_fixture, load_npz

PASSES = [("w_bob", True, ["[Bob]"], 3), ("w_alice", True, ["[Alice]"], 2),
          ("r_bob", False, ["[Bob]"], 1), ("r_both", False, ["[Bob]", "[Alice]"], 1)]


def scenario():
    z = load_npz("lowvram.npz")
    H, W, C, heads, Fl, steps, n_layers = (int(x) for x in z["params"])
    return z, H, W, C, heads, Fl, steps, n_layers


def replay(make_procs, set_state, dtype=torch.float32, device="cpu", seed=2047):
    """make_procs(n_layers, Fl) -> list of processors; set_state(**globals) updates the control globals.
    Yields (tag, step, layer, got, want, draws_got, draws_want)."""
    z, H, W, C, heads, Fl, steps, n_layers = scenario()
    attns = [attn_from_fixture(z, f"w{li}_", C, heads, dtype=dtype, device=device) for li in range(n_layers)]
    torch.manual_seed(seed)
    if device != "cpu":
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    # the golden run created the FakeAttention weights from the torch stream BEFORE anything was sampled: consume the
    # same amount so that the sampler sees the same generator state
    from oracle.fake_diffusers import FakeAttention
    for _ in range(n_layers):
        FakeAttention(C, heads)
    set_state(height=H, width=W, sa32=0.5, sa64=0.5, total_count=n_layers)
    procs = make_procs(n_layers, Fl)
    draws = []
    real = random.random

    def traced():
        u = real()
        draws.append(u)
        return u

    random.random = traced
    try:
        with torch.no_grad():
            for tag, write, chars, imgs in PASSES:
                set_state(write=write, cur_step=0, attn_count=0, cur_character=list(chars))
                for step in range(steps):
                    for li in range(n_layers):
                        x = torch.from_numpy(z[f"{tag}_s{step}_l{li}_x"]).to(device=device, dtype=dtype)
                        nb = len(draws)
                        y = procs[li](attns[li], x)
                        yield (tag, step, li, y, torch.from_numpy(z[f"{tag}_s{step}_l{li}_y"]), draws[nb:],
                               z[f"{tag}_s{step}_l{li}_draw"].tolist(), procs, z)
    finally:
        random.random = real
