"""Generate tests/golden/lowvram.npz by running the UNMODIFIED low-VRAM reference processor on CPU.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden_lowvram.py
                                                         python tests/golden/make_golden_lowvram.py --blob-only
(--blob-only leaves lowvram.npz alone and writes lowvram_bank_bob.pt: the id_bank of character "[Bob]" after the same
scenario, saved by the reference's OWN save_single_character_weights, :437-457, also executed verbatim.)

The reference class lives in StoryDiffusion/gradio_app_sdxl_specific_id_low_vram.py:99-366, a gradio application whose
module scope loads models and builds a UI — it cannot be imported.  The class definition is therefore taken VERBATIM
from the file (its top-level source extent) and executed in a namespace that provides what the class
reads at module scope: ``torch``, ``F``, ``random``, ``device``, the control globals (:548-561) and the reference's own
sampler ``cal_attn_indice_xl_effcient_memory`` (imported from StoryDiffusion/utils/gradio_utils.py through
oracle/ref_loader.py).  Nothing of the class is restated here.

Scenario (3 layers: two /32-class, one /16-class, total_count = 3; H = W = 96 (9 and 36 tokens per frame), C = 64, 1 head, id_length = 3):
  write pass of character "[Bob]"   with 3 reference images (batch 6), 4 steps
  write pass of character "[Alice]" with 2 reference images (batch 4), 4 steps
  read passes (batch 2, 4 steps each) of a frame with ["[Bob]"] and of a frame with ["[Bob]", "[Alice]"]
Recorded: every call's input and output, the gate draws, the branch taken (from the draw), the sampled index lists
of every step, and the bank contents of one layer.
"""
from __future__ import annotations

import ast
import copy
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.fake_diffusers import FakeAttention  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(ref_loader.REFERENCE_ROOT, "StoryDiffusion", "gradio_app_sdxl_specific_id_low_vram.py")


def load_reference_class():
    """Returns (namespace, class): the low-VRAM SpatialAttnProcessor2_0 exactly as written in the reference file."""
    ref_loader.load_reference()   # installs the import stubs and puts the reference root on sys.path
    gu = sys.modules["StoryDiffusion.utils.gradio_utils"]
    # the file as a whole does not parse (line 30 reads `from diffusers..utils.loading_utils import ...`), so the
    # class is cut out by its top-level extent: from its `class` line to the next top-level statement
    lines = open(REF_FILE).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.startswith("class SpatialAttnProcessor2_0("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0].isspace()
               and not lines[i].startswith("#"))
    seg = "\n".join(lines[start:end])
    node = ast.parse(seg).body[0]
    assert isinstance(node, ast.ClassDef) and node.name == "SpatialAttnProcessor2_0"
    ns = {
        "torch": torch, "F": F, "random": random, "device": "cpu",
        "cal_attn_indice_xl_effcient_memory": gu.cal_attn_indice_xl_effcient_memory,
        # control globals, :548-561
        "attn_count": 0, "total_count": 0, "cur_step": 0, "id_length": 4, "total_length": 5, "write": False,
        "sa32": 0.5, "sa64": 0.5, "height": 768, "width": 768, "indices1024": None, "indices4096": None,
        "cur_character": [], "character_dict": {}, "character_index_dict": {}, "invert_character_index_dict": {},
        "ref_indexs_dict": {}, "ref_totals": [],
    }
    exec(compile(seg, REF_FILE, "exec"), ns)
    return ns, ns["SpatialAttnProcessor2_0"]


def load_reference_function(ns, name):
    """Executes the top-level function `name` of the reference file, verbatim, in the namespace of the class."""
    lines = open(REF_FILE).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.startswith(f"def {name}("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0].isspace()
               and not lines[i].startswith("#"))
    exec(compile("\n".join(lines[start:end]), REF_FILE, "exec"), ns)
    return ns[name]


def main(blob_only=False):
    ns, cls = load_reference_class()
    H = W = 96
    C, heads, Fl, steps = 64, 1, 3, 4
    n32, n16 = (H // 32) * (W // 32), (H // 16) * (W // 16)
    layer_tokens = [n32, n32, n16]
    torch.manual_seed(2047)
    np.random.seed(2047)
    random.seed(2047)
    attns = [FakeAttention(C, heads) for _ in layer_tokens]
    g = torch.Generator().manual_seed(4321)
    ns.update(height=H, width=W, sa32=0.5, sa64=0.5, total_count=len(layer_tokens), id_length=Fl,
              total_length=Fl + 1)
    procs = copy.deepcopy([cls(id_length=Fl, device="cpu", dtype=torch.float32) for _ in layer_tokens])

    out = {"params": np.array([H, W, C, heads, Fl, steps, len(layer_tokens)], dtype=np.int64)}
    for li, a in enumerate(attns):
        for name, p in a.state_dict().items():
            out[f"w{li}_{name}"] = p.numpy()

    draws = []
    real_random = random.random

    def traced_random():
        u = real_random()
        draws.append(u)
        return u

    random.random = traced_random
    passes = [("w_bob", True, ["[Bob]"], 3), ("w_alice", True, ["[Alice]"], 2),
              ("r_bob", False, ["[Bob]"], 1), ("r_both", False, ["[Bob]", "[Alice]"], 1)]
    try:
        with torch.no_grad():
            for tag, write, chars, imgs in passes:
                ns.update(write=write, cur_step=0, attn_count=0, cur_character=list(chars))
                for step in range(steps):
                    for li, (a, p, n) in enumerate(zip(attns, procs, layer_tokens)):
                        x = torch.randn(2 * imgs, n, C, generator=g)
                        n_before = len(draws)
                        y = p(a, x.clone())
                        out[f"{tag}_s{step}_l{li}_x"] = x.numpy()
                        out[f"{tag}_s{step}_l{li}_y"] = y.numpy()
                        out[f"{tag}_s{step}_l{li}_draw"] = np.array(draws[n_before:], dtype=np.float64)
                    # index lists in force for the NEXT step (re-sampled by the last layer of this step)
                    for r, name in ((ns["indices1024"], "i32"), (ns["indices4096"], "i16")):
                        for f, ix in enumerate(r):
                            out[f"{tag}_s{step}_{name}_{f}"] = ix.numpy().astype(np.int32)
    finally:
        random.random = real_random
    if blob_only:
        import types
        save = load_reference_function(ns, "save_single_character_weights")
        unet = types.SimpleNamespace(attn_processors={f"layer{li}": p for li, p in enumerate(procs)})
        save(unet, "[Bob]", "a man, wearing a black suit", os.path.join(OUT, "lowvram_bank_bob.pt"))
        print("lowvram_bank_bob.pt written by the reference's save_single_character_weights")
        return
    out["draws"] = np.array(draws, dtype=np.float64)
    # bank of layer 2 (/16-class): per character, per step, per image (2, K, C)
    for ch, key in (("[Bob]", "bob"), ("[Alice]", "alice")):
        for step, arr in procs[2].id_bank[ch].items():
            for i, t in enumerate(arr):
                out[f"bank_{key}_s{step}_i{i}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "lowvram.npz"), **out)
    print("lowvram.npz:", len(out), "arrays,", len(draws), "gate draws")


if __name__ == "__main__":
    main(blob_only="--blob-only" in sys.argv)
