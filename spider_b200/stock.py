"""The reference's stock SDPA processor on the B200 kernels.

``AttnProcessor2_0`` of StoryDiffusion/utils/gradio_utils.py:387-472 is what the reference installs on every
attention layer that does NOT get the consistent processor — all cross-attention layers and the self-attention
layers outside ``up_blocks`` (Comic_Generation.py:353-371 instantiates it under the name ``AttnProcessor``, :17).
Its attention call (:444-446) is a plain softmax(q k^T / sqrt(d)) v per batch element: the contiguous-segment mode of
``csa_attn_fwd`` (the same launch ``__call2__`` uses), with the keys of batch element b being rows
``[b * N_k, (b + 1) * N_k)`` of the K/V projections — N_k = N for self-attention, the text length (77) for
cross-attention (a ragged 128-key tile, masked by the kernel).  Projections run on ``csa_gemm`` where the module is
a plain ``nn.Linear`` of the activations' dtype (cross-attention K/V: K = 2048), else through the module.

Same constructor and call signature as the reference class; ``attention_mask`` must be None (SDXL never passes one;
the consistent processor has the same narrowing)."""
from __future__ import annotations

import torch

from . import native


class AttnProcessor2_0(torch.nn.Module):
    native_projections = True   # csa_gemm where the shapes allow; False: always the module's own nn.Linear layers

    def __init__(self, hidden_size=None, cross_attention_dim=None):    # gradio_utils.py:391-398
        super().__init__()

    def _linear(self, lin, x2):
        """y = lin(x2) for a 2-D x2: the hand-written GEMM when lin is a plain Linear of x2's dtype and device."""
        w = getattr(lin, "weight", None)
        if (self.native_projections and type(lin) is torch.nn.Linear and w.dtype == x2.dtype
                and w.device == x2.device and w.is_contiguous() and x2.is_contiguous()
                and (lin.bias is None or (lin.bias.dtype == x2.dtype and lin.bias.is_contiguous()))
                and native.gemm_supported(x2.shape[0], w.shape[0], w.shape[1])):
            return native.gemm(x2, w.detach(), None if lin.bias is None else lin.bias.detach())
        return lin(x2)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if attention_mask is not None:
            raise NotImplementedError("AttnProcessor2_0 (B200): a caller-supplied attention_mask is not supported "
                                      "(SDXL passes none)")
        if not hidden_states.is_cuda:
            raise native.CsaNativeError("AttnProcessor2_0 (B200) got CPU hidden_states: this path has no CPU fallback")
        if hidden_states.dtype not in (torch.float16, torch.bfloat16):
            raise native.CsaNativeError(f"AttnProcessor2_0 (B200) computes in fp16/bf16; got {hidden_states.dtype}")
        residual = hidden_states                                                    # :408
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        input_ndim = hidden_states.ndim
        if input_ndim == 4:                                                         # :415-417
            b4, c4, h4, w4 = hidden_states.shape
            hidden_states = hidden_states.view(b4, c4, h4 * w4).transpose(1, 2)
        if attn.group_norm is not None:                                             # :429-430
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        x = hidden_states.contiguous()
        B, N, _ = x.shape
        if encoder_hidden_states is None:                                           # :434-437
            enc = x
        else:
            enc = encoder_hidden_states
            if attn.norm_cross:
                enc = attn.norm_encoder_hidden_states(enc)
            enc = enc.contiguous()
            if enc.shape[0] != B:
                raise ValueError(f"encoder_hidden_states has batch {enc.shape[0]}, hidden_states {B}")
        Nk = enc.shape[1]
        q = self._linear(attn.to_q, x.view(B * N, -1))                              # :432
        k = self._linear(attn.to_k, enc.view(B * Nk, -1))                           # :439-440
        v = self._linear(attn.to_v, enc.view(B * Nk, -1))
        C = q.shape[1]
        heads = attn.heads
        if C != heads * native.CSA_HEAD_DIM or k.shape[1] != C:
            raise native.CsaNativeError(f"head_dim {C // heads} != 64: not an SDXL attention layer")
        o = torch.empty_like(q)
        # softmax(q k^T / sqrt(d)) v per batch element (:444-446): keys = rows [b*Nk, (b+1)*Nk) of K / V
        native.attn_fwd(q.contiguous(), o, heads=heads, n_groups=1, n_frames=B, n_q=N,
                        k_b=k.contiguous(), v_b=v.contiguous(), b_group_rows=B * Nk, cb=(0, Nk, Nk))
        out = self._linear(attn.to_out[0], o)                                       # :452
        out = attn.to_out[1](out.view(B, N, -1))                                    # :454
        if input_ndim == 4:                                                         # :456-457
            out = out.transpose(-1, -2).reshape(b4, c4, h4, w4)
        if attn.residual_connection:
            out = out + residual
        return out / attn.rescale_output_factor                                    # :462
