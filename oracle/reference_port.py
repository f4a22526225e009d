"""CPU restatement of the reference's Consistent Self-Attention path — TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).  Every function cites the reference lines it follows (paths relative to /root/reference).

The reference keeps its control state in module globals of ``StoryDiffusion/Comic_Generation.py``
(``write, cur_step, attn_count, total_count, sa32, sa64, height, width, mask1024, mask4096``; :82-85, set by the
driver at :327-349, :372-376, :435-448).  Here that state lives on a ``StoryState`` object so that several
oracles can coexist in one test process.

The arithmetic is the same torch library calls the reference makes (``nn.Linear`` through the ``attn`` object,
``F.scaled_dot_product_attention`` with a dense boolean mask, ``torch.rand``), evaluated on CPU, normally fp32.
Pinned against the unmodified reference by tests/golden (fixtures produced by tests/golden/make_golden.py).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------
# mask sampling — StoryDiffusion/utils/gradio_utils.py:241-287 (cal_attn_mask_xl)
# ----------------------------------------------------------------------------------------------------------------
def token_counts(height: int, width: int) -> Tuple[int, int]:
    """gradio_utils.py:250-251 — tokens per frame at the /32 and /16 resolutions."""
    return (height // 32) * (width // 32), (height // 16) * (width // 16)


def sample_vectors(total_length: int, sa32: float, sa64: float, height: int, width: int, device="cpu",
                   dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """gradio_utils.py:257-258 — ONE Bernoulli(sa) row per resolution, /32 first then /16.  The two torch.rand
    calls (shape (1, T*n), given dtype/device, default generator) are the RNG contract of the path."""
    n32, n16 = token_counts(height, width)
    r32 = torch.rand((1, total_length * n32), device=device, dtype=dtype) < sa32
    r16 = torch.rand((1, total_length * n16), device=device, dtype=dtype) < sa64
    return r32[0], r16[0]


def frame_rows(sample: torch.Tensor, total_length: int, id_length: int) -> torch.Tensor:
    """gradio_utils.py:260-278 — the T distinct mask rows: row i = sample restricted to columns < id_length*n,
    with the own block [i*n, (i+1)*n) forced True.  Returns bool (T, T*n)."""
    n = sample.numel() // total_length
    rows = sample.unsqueeze(0).repeat(total_length, 1)
    for i in range(total_length):
        rows[i, id_length * n:] = False
        rows[i, i * n:(i + 1) * n] = True
    return rows


def dense_mask(rows: torch.Tensor) -> torch.Tensor:
    """gradio_utils.py:285-286 — every row repeated n times: (T*n, T*n) bool."""
    T = rows.shape[0]
    n = rows.shape[1] // T
    return rows.unsqueeze(1).repeat(1, n, 1).reshape(-1, T * n)


def cal_attn_mask_xl(total_length, id_length, sa32, sa64, height, width, device="cpu", dtype=torch.float32):
    """Same signature and RNG consumption as the reference function; returns (mask1024, mask4096)."""
    r32, r16 = sample_vectors(total_length, sa32, sa64, height, width, device, dtype)
    return dense_mask(frame_rows(r32, total_length, id_length)), dense_mask(frame_rows(r16, total_length, id_length))


def index_lists(rows: torch.Tensor) -> List[torch.Tensor]:
    """What the CUDA compaction must reproduce bit-exactly: ascending attended columns of each distinct row
    (== torch.nonzero(mask[f*n]) of the dense mask)."""
    return [torch.nonzero(rows[i], as_tuple=True)[0].to(torch.int32) for i in range(rows.shape[0])]


# ----------------------------------------------------------------------------------------------------------------
# state + processor — StoryDiffusion/Comic_Generation.py:46-268
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class StoryState:
    """The module globals of Comic_Generation.py (:82-85) as an object."""

    write: bool = False
    cur_step: int = 0
    attn_count: int = 0
    total_count: int = 0
    sa32: float = 0.5
    sa64: float = 0.5
    height: int = 768
    width: int = 768
    mask1024: Optional[torch.Tensor] = None
    mask4096: Optional[torch.Tensor] = None
    trace: List[tuple] = field(default_factory=list)  # (branch, cur_step, random draw or None), for KATs


def _project_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = x.shape
    return x.view(b, n, heads, c // heads).transpose(1, 2)


class ConsistentAttnOracle(torch.nn.Module):
    """Restatement of ``SpatialAttnProcessor2_0`` (Comic_Generation.py:46-268) bound to a ``StoryState``."""

    def __init__(self, state: StoryState, hidden_size=None, cross_attention_dim=None, id_length=4, device="cpu",
                 dtype=torch.float32):
        super().__init__()
        self.state = state
        self.device = device          # :66  only used for mask regeneration and the bank .to()
        self.dtype = dtype            # :67
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.total_length = id_length + 1   # :70
        self.id_length = id_length          # :71
        self.id_bank: Dict[int, list] = {}  # :72

    # -- :74-127 ------------------------------------------------------------------------------------------------
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        st = self.state
        F_ = self.id_length
        if st.write:
            # :87-89 the bank aliases the layer input (views, no clone)
            self.id_bank[st.cur_step] = [hidden_states[:F_], hidden_states[F_:]]
        else:
            # :90-92 K/V source = [bank uncond frames, current uncond, bank cond frames, current cond]
            bank_u, bank_c = self.id_bank[st.cur_step]
            encoder_hidden_states = torch.cat(
                (bank_u.to(self.device), hidden_states[:1], bank_c.to(self.device), hidden_states[1:]))
        if st.cur_step < 5:
            # :94-96 early steps: standard attention (over bank + self when reading)
            st.trace.append(("standard-early", st.cur_step, None))
            out = self.standard(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
        else:
            u = random.random()                              # :98   one draw per call
            thr = 0.3 if st.cur_step < 20 else 0.1           # :99-102
            if u > thr:
                n = hidden_states.shape[1]
                use32 = n == (st.height // 32) * (st.width // 32)   # :105 / :110
                mask = st.mask1024 if use32 else st.mask4096
                cut = mask.shape[0] // self.total_length * self.id_length
                if not st.write:
                    attention_mask = mask[cut:]              # :106 / :108  rows of the new frame, all columns
                else:
                    attention_mask = mask[:cut, :cut]        # :111 / :113
                st.trace.append(("consistent", st.cur_step, u))
                out = self.consistent(attn, hidden_states, encoder_hidden_states, attention_mask, temb)
            else:
                st.trace.append(("standard", st.cur_step, u))
                out = self.standard(attn, hidden_states, None, attention_mask, temb)   # :118  bank ignored
        st.attn_count += 1                                   # :119
        if st.attn_count == st.total_count:                  # :120-125
            st.attn_count = 0
            st.cur_step += 1
            st.mask1024, st.mask4096 = cal_attn_mask_xl(self.total_length, self.id_length, st.sa32, st.sa64,
                                                        st.height, st.width, device=self.device, dtype=self.dtype)
        return out

    # -- :129-196 (__call1__) -----------------------------------------------------------------------------------
    def consistent(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        four_d = hidden_states.ndim == 4
        if four_d:
            tb, ch, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(tb, ch, hh * ww).transpose(1, 2)
        total_batch, n_tok, ch = hidden_states.shape
        frames = total_batch // 2                                        # :146
        # :148 fold the frames of each CFG half into one sequence
        hidden_states = hidden_states.view(-1, frames, n_tok, ch).reshape(-1, frames * n_tok, ch)
        batch = hidden_states.shape[0]
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = attn.to_q(hidden_states)                                     # :155
        if encoder_hidden_states is None:
            kv_src = hidden_states                                       # :159
        else:
            kv_src = encoder_hidden_states.view(-1, self.id_length + 1, n_tok, ch).reshape(
                -1, (self.id_length + 1) * n_tok, ch)                    # :162
        k = attn.to_k(kv_src)                                            # :164
        v = attn.to_v(kv_src)                                            # :165
        q, k, v = (_project_heads(t, attn.heads) for t in (q, k, v))     # :171-174
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(total_batch, -1, ch).to(q.dtype)   # :179-180
        o = attn.to_out[1](attn.to_out[0](o))                            # :185-187
        if four_d:
            o = o.transpose(-1, -2).reshape(total_batch, ch, hh, ww)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor                            # :194

    # -- :198-268 (__call2__) -----------------------------------------------------------------------------------
    def standard(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        four_d = hidden_states.ndim == 4
        if four_d:
            b0, ch, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b0, ch, hh * ww).transpose(1, 2)
        batch, n_tok, ch = hidden_states.shape
        if attention_mask is not None:                                   # :221-225
            attention_mask = attn.prepare_attention_mask(attention_mask, n_tok, batch)
            attention_mask = attention_mask.view(batch, attn.heads, -1, attention_mask.shape[-1])
        if attn.group_norm is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = attn.to_q(hidden_states)                                     # :230
        if encoder_hidden_states is None:
            kv_src = hidden_states                                       # :233
        else:
            kv_src = encoder_hidden_states.view(-1, self.id_length + 1, n_tok, ch).reshape(
                -1, (self.id_length + 1) * n_tok, ch)                    # :235
        k = attn.to_k(kv_src)
        v = attn.to_v(kv_src)
        q, k, v = (_project_heads(t, attn.heads) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(batch, -1, ch).to(q.dtype)         # :252-253
        o = attn.to_out[1](attn.to_out[0](o))
        if four_d:
            o = o.transpose(-1, -2).reshape(batch, ch, hh, ww)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor                            # :266


# ----------------------------------------------------------------------------------------------------------------
# gathered form — for shapes whose dense mask cannot be materialised (F >= 16: 4.85 GB bool)
# ----------------------------------------------------------------------------------------------------------------
def gathered_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, key_lists: List[torch.Tensor],
                       heads: int) -> torch.Tensor:
    """softmax(Q K^T / sqrt(d) + M) V of Comic_Generation.py:175-177 restated per frame over index lists.

    q: (B, F*N, C) folded queries; k, v: (B, Nk, C); key_lists[f]: ascending key positions attended by every query
    of frame f (M's row block f).  Equal to the dense-masked form up to fp32 rounding (3e-7 measured).  fp32.
    """
    B, FN, C = q.shape
    Fq = len(key_lists)
    N = FN // Fq
    out = torch.empty((B, FN, C), dtype=torch.float32)
    for f, keys in enumerate(key_lists):
        keys = keys.long()
        qf = _project_heads(q[:, f * N:(f + 1) * N].float(), heads)
        kf = _project_heads(k[:, keys].float(), heads)
        vf = _project_heads(v[:, keys].float(), heads)
        of = F.scaled_dot_product_attention(qf, kf, vf)
        out[:, f * N:(f + 1) * N] = of.transpose(1, 2).reshape(B, N, C)
    return out


def setup_seed(seed: int) -> None:
    """Comic_Generation.py:35-40 (cudnn flag omitted: CPU oracle)."""
    import numpy as np

    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


# ----------------------------------------------------------------------------------------------------------------
# the stock processor of every other attention layer
# ----------------------------------------------------------------------------------------------------------------
def stock_attention(attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                    temb=None) -> torch.Tensor:
    """``AttnProcessor2_0.__call__`` of StoryDiffusion/utils/gradio_utils.py:400-464 with ``attention_mask=None``
    (what SDXL passes): optional spatial / group norm (:410-411, :429-430), 4-D input folded to tokens (:415-417),
    q from the hidden states and K/V from the encoder states — the hidden states themselves for self-attention,
    optionally normed for cross-attention (:432-440) —, per-head softmax(q k^T / sqrt(d)) v (:444-446), output
    projection + dropout (:452-454), un-folding, residual and rescale (:456-462).  fp32.  Pinned to the unmodified
    reference class by tests/golden/stock.npz (tests/golden/make_golden_stock.py)."""
    residual = hidden_states
    if attn.spatial_norm is not None:
        hidden_states = attn.spatial_norm(hidden_states, temb)
    ndim = hidden_states.ndim
    if ndim == 4:
        b, c, h, w = hidden_states.shape
        hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
    if attn.group_norm is not None:
        hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
    enc = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    if encoder_hidden_states is not None and attn.norm_cross:
        enc = attn.norm_encoder_hidden_states(enc)
    q = _project_heads(attn.to_q(hidden_states).float(), attn.heads)
    k = _project_heads(attn.to_k(enc).float(), attn.heads)
    v = _project_heads(attn.to_v(enc).float(), attn.heads)
    o = F.scaled_dot_product_attention(q, k, v)
    B, _, N, d = o.shape
    o = o.transpose(1, 2).reshape(B, N, attn.heads * d)
    out = attn.to_out[1](attn.to_out[0](o))
    if ndim == 4:
        out = out.transpose(-1, -2).reshape(b, c, h, w)
    if attn.residual_connection:
        out = out + residual
    return out / attn.rescale_output_factor
