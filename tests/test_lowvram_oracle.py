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


def test_bank_file_written_by_the_reference_loads_and_our_file_equals_it(tmp_path):
    """tests/golden/lowvram_bank_bob.pt was written by the reference's OWN save_single_character_weights
    (gradio_app_sdxl_specific_id_low_vram.py:437-457, executed verbatim by make_golden_lowvram.py --blob-only) after the
    golden story's write pass of "[Bob]".  The product loader must fill the processors' banks from it with exactly the
    tensors the reference recorded, and the product saver must write the same file back (same keys, same tensors)."""
    import os
    import types

    from spider_b200 import lowvram

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lowvram_bank_bob.pt")
    z = np.load(os.path.join(os.path.dirname(path), "lowvram.npz"))
    host = types.SimpleNamespace(attn_count=0, total_count=3, cur_step=0, id_length=3, total_length=4, write=False,
                                 sa32=0.5, sa64=0.5, height=96, width=96, indices1024=None, indices4096=None,
                                 cur_character=["[Bob]"], character_dict={}, character_index_dict={},
                                 invert_character_index_dict={}, ref_indexs_dict={}, ref_totals=[])
    cls = lowvram.make_lowvram_processor_class(host)
    unet = types.SimpleNamespace(attn_processors={f"layer{li}": cls(id_length=3, device="cpu") for li in range(3)},
                                 device=torch.device("cpu"))
    ch, desc = lowvram.load_single_character_weights(unet, path)
    assert (ch, desc) == ("[Bob]", "a man, wearing a black suit")
    bank = unet.attn_processors["layer2"].id_bank["[Bob]"]
    assert sorted(bank) == [0, 1, 2, 3]
    for step in range(4):
        assert len(bank[step]) == 3
        for i, t in enumerate(bank[step]):
            assert np.array_equal(t.numpy(), z[f"bank_bob_s{step}_i{i}"])
    # and back: the product's file holds what the reference's file holds
    out = os.path.join(tmp_path, "bob.pt")
    lowvram.save_single_character_weights(unet, "[Bob]", "a man, wearing a black suit", out)
    ours, theirs = torch.load(out, map_location="cpu"), torch.load(path, map_location="cpu")
    assert sorted(ours) == sorted(theirs)
    for name in (k for k in theirs if k.startswith("layer")):
        assert sorted(ours[name]) == sorted(theirs[name])
        for step in theirs[name]:
            assert len(ours[name][step]) == len(theirs[name][step])
            for a, b in zip(ours[name][step], theirs[name][step]):
                assert torch.equal(a, b)
