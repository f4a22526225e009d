"""id_bank: the write-once / read-many cache of the identity frames.

Reference: ``self.id_bank[cur_step] = [hidden_states[:F], hidden_states[F:]]`` in write mode
(StoryDiffusion/Comic_Generation.py:87-89, views of the layer input) and, in read mode, a ``torch.cat`` of those
with the current frame followed by a re-projection of all ``2*(F+1)*N`` rows through ``to_k``/``to_v`` on every call
(:92, :162-165).

Here the write pass keeps the *projected* K and V of the identity frames — they are computed by the write pass
anyway, so the GEMM epilogue writes straight into the bank (no copy kernel, no extra launch) — and the read pass
hands those tensors to the attention kernel as K/V source A next to the current frame's source B: no ``cat``, no
re-projection, no dense mask.

``IdBank`` stays a ``dict`` keyed by ``cur_step`` and each entry still behaves like the reference's two-element list
``[uncond, cond]`` so that code that pokes ``proc.id_bank`` (StoryDiffusion/gradio_app_sdxl_specific_id_low_vram.py
:437-479, :746-752) keeps working:  what the list holds depends on ``store``:
  "kv"      (default) entry[0]/entry[1] are None placeholders, only K/V are kept (2x the reference's memory).
  "hidden"  exactly the reference's views; K/V are projected lazily on the first read and cached on the entry.
  "both"    hidden views and K/V.
"""
from __future__ import annotations

from typing import Optional

import torch

STORE_MODES = ("kv", "hidden", "both")


class BankEntry(list):
    """``[hidden_uncond, hidden_cond]`` (reference layout) + projected ``k``/``v`` of shape ``(2*F*N, C)``."""

    def __init__(self, hidden_u: Optional[torch.Tensor] = None, hidden_c: Optional[torch.Tensor] = None,
                 k: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None):
        super().__init__([hidden_u, hidden_c])
        self.k = k
        self.v = v

    def has_kv(self) -> bool:
        return self.k is not None and self.v is not None

    def kv(self, attn, device=None):
        """Projected K/V rows ``[uncond frames..., cond frames...]``; projects (and caches) from the hidden views
        when the entry came from reference-style code (a plain ``[hs_u, hs_c]`` list or a loaded checkpoint)."""
        if not self.has_kv():
            hu, hc = self[0], self[1]
            if hu is None or hc is None:
                raise KeyError("id_bank entry holds neither K/V nor hidden states")
            if device is not None:
                hu, hc = hu.to(device), hc.to(device)          # Comic_Generation.py:92 `.to(self.device)`
            c = hu.shape[-1]
            src = torch.cat((hu.reshape(-1, c), hc.reshape(-1, c)))
            self.k = attn.to_k(src)
            self.v = attn.to_v(src)
        return self.k, self.v


class IdBank(dict):
    """dict[cur_step] -> BankEntry.  Plain ``[hs_u, hs_c]`` lists assigned from outside are wrapped on access."""

    def __getitem__(self, step):
        e = super().__getitem__(step)   # KeyError for a step never written, like the reference (:92)
        if not isinstance(e, BankEntry):
            e = BankEntry(e[0], e[1])
            super().__setitem__(step, e)
        return e

    def nbytes(self) -> int:
        tot = 0
        for e in self.values():
            for t in (list(e) + [getattr(e, "k", None), getattr(e, "v", None)]):
                if isinstance(t, torch.Tensor):
                    tot += t.numel() * t.element_size()
        return tot
