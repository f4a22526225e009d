"""CPU: the low-VRAM oracle (oracle/reference_port_lowvram.py) against the golden fixture produced by the UNMODIFIED
reference class (tests/golden/make_golden_lowvram.py).  This is what pins that oracle."""
import numpy as np
import torch

from oracle import reference_port_lowvram as rl

from lowvram_story import replay, scenario


def test_lowvram_oracle_replays_reference_story_bit_for_bit():
    st = rl.LowVramState()

    def set_state(**kw):
        for k, v in kw.items():
            setattr(st, k, v)

    def make(n_layers, Fl):
        return [rl.LowVramOracle(st, id_length=Fl, device="cpu", dtype=torch.float32) for _ in range(n_layers)]

    n = 0
    last = None
    for tag, step, li, got, want, dg, dw, procs, z in replay(make, set_state):
        assert dg == dw, f"{tag} step {step} layer {li}: gate draws differ"
        assert got.shape == want.shape
        err = (got - want).abs().max().item()
        assert err <= 1e-5, f"{tag} step {step} layer {li}: max-abs {err:.3e}"
        if li == len(procs) - 1:
            # index lists re-sampled by the step's last layer: same RNG consumption as the reference sampler
            for name, lists in (("i32", st.indices1024), ("i16", st.indices4096)):
                for f, ix in enumerate(lists):
                    assert np.array_equal(ix.numpy().astype(np.int32), z[f"{tag}_s{step}_{name}_{f}"])
        n += 1
        last = (procs, z)
    assert n == 4 * 4 * 3
    procs, z = last
    # bank layout: character -> step -> one (2, K_img, C) tensor per reference image
    for ch, key, imgs in (("[Bob]", "bob", 3), ("[Alice]", "alice", 2)):
        bank = procs[2].id_bank[ch]
        assert sorted(bank) == [0, 1, 2, 3]
        for step, arr in bank.items():
            assert len(arr) == imgs
            for i, t in enumerate(arr):
                assert np.allclose(t.numpy(), z[f"bank_{key}_s{step}_i{i}"], atol=0)


def test_lowvram_sampler_shapes_and_rng():
    torch.manual_seed(5)
    a, b = rl.cal_attn_indice_xl_effcient_memory(4, 3, 0.3, 0.7, 96, 64)
    torch.manual_seed(5)
    m32 = torch.rand((4, 6)) < 0.3
    m16 = torch.rand((4, 24)) < 0.7
    assert len(a) == len(b) == 4
    for i in range(4):
        assert torch.equal(a[i], torch.nonzero(m32[i], as_tuple=True)[0])
        assert torch.equal(b[i], torch.nonzero(m16[i], as_tuple=True)[0])
