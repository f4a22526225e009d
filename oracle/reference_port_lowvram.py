"""CPU restatement of the reference's LOW-VRAM Consistent Self-Attention variant — TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  Every function cites the reference lines it follows (paths relative to /root/reference).

Reference: ``SpatialAttnProcessor2_0`` of StoryDiffusion/gradio_app_sdxl_specific_id_low_vram.py:99-366 and the
sampler ``cal_attn_indice_xl_effcient_memory`` of StoryDiffusion/utils/gradio_utils.py:303-312.  Differences from the
main variant (oracle/reference_port.py): every frame owns an independent list of sampled token positions
(``indices1024 / indices4096``), the id_bank keeps only the sampled tokens of every reference image, per character
(:172-179), the early cutoff is ``cur_step < 1`` (:192), each image attends the other images' sampled tokens followed by
ALL of its own tokens (:231-246), and a read frame attends the bank tokens of all its characters plus itself
(:186-190, :252-261).

The control state (module globals of the gradio app, :145-149, :548-561) lives on a ``LowVramState`` object.  Pinned
against the unmodified reference class by tests/golden/lowvram.npz (tests/golden/make_golden_lowvram.py executes the
class taken verbatim from the reference file).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def cal_attn_indice_xl_effcient_memory(total_length, id_length, sa32, sa64, height, width, device="cpu",
                                       dtype=torch.float32):
    """gradio_utils.py:303-312 — one Bernoulli(sa) row PER FRAME and resolution, returned as index lists.  The two
    torch.rand calls of shape (T, n) are the RNG contract."""
    nums_1024 = (height // 32) * (width // 32)
    nums_4096 = (height // 16) * (width // 16)
    bool_matrix1024 = torch.rand((total_length, nums_1024), device=device, dtype=dtype) < sa32
    bool_matrix4096 = torch.rand((total_length, nums_4096), device=device, dtype=dtype) < sa64
    indices1024 = [torch.nonzero(bool_matrix1024[i], as_tuple=True)[0] for i in range(total_length)]
    indices4096 = [torch.nonzero(bool_matrix4096[i], as_tuple=True)[0] for i in range(total_length)]
    return indices1024, indices4096


@dataclass
class LowVramState:
    """Module globals of gradio_app_sdxl_specific_id_low_vram.py (:145-149, :548-561) as an object."""

    write: bool = False
    cur_step: int = 0
    attn_count: int = 0
    total_count: int = 0
    sa32: float = 0.5
    sa64: float = 0.5
    height: int = 768
    width: int = 768
    indices1024: Optional[list] = None
    indices4096: Optional[list] = None
    cur_character: List[str] = field(default_factory=list)
    trace: List[tuple] = field(default_factory=list)   # (branch, cur_step, random draw or None)


def _heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = x.shape
    return x.view(b, n, heads, c // heads).transpose(1, 2)


class LowVramOracle(torch.nn.Module):
    """Restatement of the low-VRAM ``SpatialAttnProcessor2_0`` (:99-366) bound to a ``LowVramState``."""

    def __init__(self, state: LowVramState, hidden_size=None, cross_attention_dim=None, id_length=4, device="cpu",
                 dtype=torch.float32):
        super().__init__()
        self.state = state
        self.device = device
        self.dtype = dtype
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.total_length = id_length + 1          # :130
        self.id_length = id_length                 # :131
        self.id_bank: Dict[str, Dict[int, list]] = {}   # :132  character -> step -> [ (2, K_img, C) per image ]

    def _resample(self):
        st = self.state
        st.indices1024, st.indices4096 = cal_attn_indice_xl_effcient_memory(
            self.total_length, self.id_length, st.sa32, st.sa64, st.height, st.width, device=self.device,
            dtype=self.dtype)

    # -- :137-282 -----------------------------------------------------------------------------------------------
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        st = self.state
        if st.attn_count == 0 and st.cur_step == 0:                      # :150-160
            self._resample()
        n32 = (st.height // 32) * (st.width // 32)
        encoder_arr = None
        if st.write:                                                     # :161-180
            assert len(st.cur_character) == 1
            indices = st.indices1024 if hidden_states.shape[1] == n32 else st.indices4096
            total_batch, n_tok, ch = hidden_states.shape
            img_nums = total_batch // 2
            hs4 = hidden_states.reshape(-1, img_nums, n_tok, ch)
            bank = self.id_bank.setdefault(st.cur_character[0], {})
            bank[st.cur_step] = [hs4[:, i, indices[i], :].reshape(2, -1, ch).clone() for i in range(img_nums)]
        else:                                                            # :181-190
            encoder_arr = []
            for character in st.cur_character:
                encoder_arr = encoder_arr + [t.to(self.device) for t in self.id_bank[character][st.cur_step]]
        if st.cur_step < 1:                                              # :192-195
            st.trace.append(("standard-early", st.cur_step, None))
            hidden_states = self.call2(attn, hidden_states, None, attention_mask, temb)
        else:
            u = random.random()                                          # :197
            thr = 0.3 if st.cur_step < 20 else 0.1                       # :198-201
            if u > thr:
                st.trace.append(("consistent", st.cur_step, u))
                indices = st.indices1024 if hidden_states.shape[1] == n32 else st.indices4096
                if st.write:                                             # :209-246
                    total_batch, n_tok, ch = hidden_states.shape
                    img_nums = total_batch // 2
                    hs4 = hidden_states.reshape(-1, img_nums, n_tok, ch).clone()
                    enc = [hs4[:, i, indices[i], :].reshape(2, -1, ch) for i in range(img_nums)]
                    out4 = hs4.clone()
                    for i in range(img_nums):
                        others = [j for j in range(img_nums) if j != i]
                        # NOTE the reference updates hidden_states in place image by image (:240-246); `enc` was
                        # gathered before the loop (advanced indexing copies), so later images still see the
                        # ORIGINAL sampled tokens of earlier images, while their own tokens are untouched until
                        # their turn.  out4 / hs4 keep the two apart.
                        tmp = torch.cat([enc[j] for j in others] + [hs4[:, i]], dim=1)
                        out4[:, i] = self.call2(attn, hs4[:, i], tmp, None, temb)
                    hidden_states = out4.reshape(-1, n_tok, ch)
                else:                                                    # :247-262
                    _, n_tok, ch = hidden_states.shape
                    hs4 = hidden_states.reshape(2, -1, n_tok, ch).clone()
                    tmp = torch.cat(encoder_arr + [hs4[:, 0]], dim=1)
                    hs4[:, 0] = self.call2(attn, hs4[:, 0], tmp, None, temb)
                    hidden_states = hs4.reshape(-1, n_tok, ch)
            else:
                st.trace.append(("standard", st.cur_step, u))
                hidden_states = self.call2(attn, hidden_states, None, attention_mask, temb)   # :263-266
        st.attn_count += 1                                               # :267-280
        if st.attn_count == st.total_count:
            st.attn_count = 0
            st.cur_step += 1
            self._resample()
        return hidden_states

    # -- :284-366 (__call2__) -----------------------------------------------------------------------------------
    def call2(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        four_d = hidden_states.ndim == 4
        if four_d:
            b0, ch, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b0, ch, hh * ww).transpose(1, 2)
        batch, n_tok, ch = hidden_states.shape
        if attention_mask is not None:
            attention_mask = attn.prepare_attention_mask(attention_mask, n_tok, batch)
            attention_mask = attention_mask.view(batch, attn.heads, -1, attention_mask.shape[-1])
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = attn.to_q(hidden_states)                                     # :318
        kv_src = hidden_states if encoder_hidden_states is None else encoder_hidden_states   # :320-323
        k = attn.to_k(kv_src)                                            # :325
        v = attn.to_v(kv_src)                                            # :326
        q, k, v = (_heads(t, attn.heads) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, ch).to(q.dtype)         # :341-344
        o = attn.to_out[1](attn.to_out[0](o))                            # :347-349
        if four_d:
            o = o.transpose(-1, -2).reshape(batch, ch, hh, ww)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor                            # :359
