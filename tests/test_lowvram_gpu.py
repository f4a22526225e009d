"""GPU: the low-VRAM processor (spider_b200/lowvram.py) through the C ABI against the golden story of the UNMODIFIED
reference class (tests/golden/lowvram.npz) and against the CPU oracle at SDXL sizes; persistence round trip in the
reference's on-disk format."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import reference_port_lowvram as rl
from oracle.fake_diffusers import FakeAttention, FakeUNet
from spider_b200 import lowvram, native
from spider_b200.processor import StoryGlobals

from helpers import MAX_ABS, MIN_COS, max_abs_cos
from lowvram_story import replay

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _host():
    h = StoryGlobals()
    h.indices1024 = h.indices4096 = None
    h.cur_character = []
    return h


@pytest.mark.parametrize("bank_store", ["hidden", "kv"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_lowvram_story_vs_reference_golden(bank_store, dtype):
    """Two characters written (3 and 2 reference images), two frames read (one and two characters), 4 steps x 3 layers:
    same seeds, same inputs, same gate draws, same index lists (sampled on the CPU generator like the golden run) —
    every call's output against the reference's."""
    host = _host()
    cls = lowvram.make_lowvram_processor_class(host, bank_store=bank_store)

    def set_state(**kw):
        for k, v in kw.items():
            setattr(host, k, v)

    def make(n_layers, Fl):
        return [cls(id_length=Fl, device="cpu", dtype=torch.float32) for _ in range(n_layers)]

    worst = (0.0, 1.0)
    n = 0
    last = None
    for tag, step, li, got, want, dg, dw, procs, z in replay(make, set_state, dtype=dtype, device=DEV):
        assert dg == dw, f"{tag} step {step} layer {li}: gate draws differ"
        err, cos = max_abs_cos(got, want)
        assert err <= MAX_ABS and cos >= MIN_COS, f"{tag} s{step} l{li}: max-abs {err:.3e} cos {cos:.6f}"
        worst = (max(worst[0], err), min(worst[1], cos))
        if li == len(procs) - 1:
            for name, ind in (("i32", host.indices1024), ("i16", host.indices4096)):
                for f in range(len(ind)):
                    assert np.array_equal(ind[f].cpu().numpy().astype(np.int32), z[f"{tag}_s{step}_{name}_{f}"])
        n += 1
        last = (procs, z)
    assert n == 48
    procs, z = last
    for ch, key, imgs in (("[Bob]", "bob", 3), ("[Alice]", "alice", 2)):
        bank = procs[2].id_bank[ch]
        assert sorted(bank) == [0, 1, 2, 3]
        for step, entry in bank.items():
            assert len(entry) == imgs
            for i, t in enumerate(entry):
                want = torch.from_numpy(z[f"bank_{key}_s{step}_i{i}"])
                assert tuple(t.shape) == tuple(want.shape)
                assert torch.equal(t.cpu(), want.to(dtype))      # sampled rows of the (rounded) layer input, bit-exact
    print(f"low-VRAM story parity ({bank_store}, {dtype}): worst max-abs {worst[0]:.3e}, worst cos {worst[1]:.6f}")


def test_lowvram_vs_oracle_sdxl_layer():
    """A 32x32-class SDXL layer (1024 tokens, 640 channels, 10 heads — half width to keep the CPU oracle quick):
    write with 4 reference images then a read frame, consistent branch, bf16 on the GPU vs the fp32 oracle."""
    H = W = 1024
    C, heads, Fl, imgs = 640, 10, 4, 4
    N = (H // 32) * (W // 32)
    torch.manual_seed(3)
    attn = FakeAttention(C, heads)
    xw = torch.randn(2 * imgs, N, C)
    xr = torch.randn(2, N, C)
    st = rl.LowVramState(total_count=10 ** 9, height=H, width=W, cur_character=["[A]"])
    orc = rl.LowVramOracle(st, id_length=Fl)
    host = _host()
    host.height, host.width, host.total_count, host.cur_character = H, W, 10 ** 9, ["[A]"]
    cls = lowvram.make_lowvram_processor_class(host)
    proc = cls(id_length=Fl, device="cpu", dtype=torch.float32)
    gattn = FakeAttention(C, heads)
    gattn.load_state_dict(attn.state_dict())
    gattn = gattn.to(DEV, torch.bfloat16)
    with torch.no_grad():
        for write, x in ((True, xw), (False, xr)):
            st.write = host.write = write
            st.cur_step = host.cur_step = 25
            st.attn_count = host.attn_count = 1          # not the first call of a pass: keep the index lists
            if write:
                torch.manual_seed(11)
                orc._resample()
                torch.manual_seed(11)
                proc._resample(host)
            random.seed(1)
            want = orc(attn, x)
            random.seed(1)
            got = proc(gattn, x.to(DEV, torch.bfloat16))
            assert st.trace[-1][0] == "consistent" and proc._last_branch == "consistent"
            err, cos = max_abs_cos(got, want)
            assert err <= MAX_ABS and cos >= MIN_COS, f"write={write}: max-abs {err:.3e} cos {cos:.6f}"


def test_lowvram_persistence_round_trip(tmp_path):
    """save_single_character_weights writes the reference's format ({description, character, name: {step: [cpu (2, K,
    C) tensors]}}); loading it back (or into a fresh processor) reproduces the read output."""
    H = W = 256
    C, heads, Fl = 128, 2, 3
    N = (H // 32) * (W // 32)
    layout = {"up_blocks.0.attentions.0.transformer_blocks.0": (C, heads)}
    host = _host()
    host.height, host.width, host.cur_character = H, W, ["[A]"]
    cls = lowvram.make_lowvram_processor_class(host)
    torch.manual_seed(0)
    random.seed(0)

    def build():
        unet = FakeUNet(layout, dtype=torch.float16, device=DEV)
        unet.device = torch.device(DEV)
        unet.set_attn_processor({n: cls(id_length=Fl, device=DEV) for n in unet.attn_processors})
        return unet

    unet = build()
    (name, attn), = unet.self_attn_layers()
    host.total_count = 1
    xw = torch.randn(2 * Fl, N, C, device=DEV, dtype=torch.float16)
    xr = torch.randn(2, N, C, device=DEV, dtype=torch.float16)
    with torch.no_grad():
        host.write, host.cur_step, host.attn_count = True, 0, 0
        for _ in range(3):
            attn(xw)                                        # steps 0, 1, 2 written
        path = os.path.join(tmp_path, "A.pt")
        lowvram.save_single_character_weights(unet, "[A]", "a person", path)
        blob = torch.load(path, map_location="cpu")
        assert blob["character"] == "[A]" and blob["description"] == "a person"
        steps = blob[name]
        assert sorted(steps) == [0, 1, 2]
        for arr in steps.values():
            assert len(arr) == Fl and all(t.device.type == "cpu" and t.shape[0] == 2 and t.shape[2] == C for t in arr)
        host.write, host.cur_step, host.attn_count = False, 1, 0
        st = random.getstate()
        want = attn(xr)
        # a fresh UNet with the same weights, bank loaded from the file
        unet2 = build()
        unet2.load_state_dict(unet.state_dict())
        ch, desc = lowvram.load_single_character_weights(unet2, path)
        assert (ch, desc) == ("[A]", "a person")
        (_, attn2), = unet2.self_attn_layers()
        host.cur_step, host.attn_count = 1, 0
        random.setstate(st)
        got = attn2(xr)
    assert torch.equal(got, want)
