"""Stand-ins for the pieces of diffusers (EXTERNAL, diffusers==0.25.0, not installed here) that the reference
processor touches: the ``Attention`` module object (attribute reads listed in SURVEY.md §8b, used at
StoryDiffusion/Comic_Generation.py:138-194, :205-266) and the ``unet.attn_processors`` /
``unet.set_attn_processor`` plugin surface (Comic_Generation.py:353-371).  Test infrastructure only.
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn


class FakeAttention(nn.Module):
    """The members of diffusers' ``Attention`` that ``SpatialAttnProcessor2_0`` reads (SDXL self-attention:
    bias-free q/k/v projections, biased output projection, dropout 0, no norms, no residual)."""

    def __init__(self, channels: int, heads: int, processor=None, dtype=torch.float32, device="cpu"):
        super().__init__()
        assert channels % heads == 0
        self.heads = heads
        self.inner_dim = channels
        self.scale = (channels // heads) ** -0.5
        self.to_q = nn.Linear(channels, channels, bias=False)
        self.to_k = nn.Linear(channels, channels, bias=False)
        self.to_v = nn.Linear(channels, channels, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = processor
        self.to(device=device, dtype=dtype)
        self.eval()

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        if attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        return attention_mask

    def set_processor(self, processor):
        # diffusers: a processor that is an nn.Module is registered as a submodule (so .to()/deepcopy see it)
        if isinstance(processor, nn.Module):
            self._modules.pop("processor", None)
            self.processor = processor
        else:
            self.__dict__["processor"] = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FakeUNet(nn.Module):
    """An SDXL-*shaped* set of self-attention (attn1) and cross-attention (attn2) layers with diffusers' naming.

    ``layout`` maps a name prefix such as ``up_blocks.0.attentions.1.transformer_blocks.3`` to (channels, heads).
    ``sdxl_layout()`` gives the full SDXL-base attn1 placement (EXTERNAL config): down 4 @640
    + 20 @1280, mid 10 @1280, up 30 @1280 + 6 @640 = 70 self-attention layers.
    """

    def __init__(self, layout: Dict[str, tuple], dtype=torch.float32, device="cpu", with_cross=False):
        super().__init__()
        self.config = _Cfg(block_out_channels=[320, 640, 1280], cross_attention_dim=2048)
        self.attn = nn.ModuleDict()
        self._names = []
        for prefix, (ch, heads) in layout.items():
            key = prefix.replace(".", "_")
            self.attn[key + "_attn1"] = FakeAttention(ch, heads, dtype=dtype, device=device)
            self._names.append((prefix + ".attn1.processor", key + "_attn1"))
            if with_cross:
                self.attn[key + "_attn2"] = FakeAttention(ch, heads, dtype=dtype, device=device)
                self._names.append((prefix + ".attn2.processor", key + "_attn2"))

    @property
    def attn_processors(self):
        return {name: self.attn[key].processor for name, key in self._names}

    def set_attn_processor(self, processors):
        if isinstance(processors, dict):
            if set(processors.keys()) != {n for n, _ in self._names}:
                raise ValueError("A dict of processors was passed, but it does not cover every attention layer")
            for name, key in self._names:
                self.attn[key].set_processor(processors[name])
        else:
            for _, key in self._names:
                self.attn[key].set_processor(processors)

    def self_attn_layers(self):
        """[(name, FakeAttention)] for attn1 layers in UNet execution order (down, mid, up)."""
        return [(n, self.attn[k]) for n, k in self._names if n.endswith("attn1.processor")]


def sdxl_layout(up_only: bool = False) -> Dict[str, tuple]:
    """Self-attention placement of stabilityai/stable-diffusion-xl-base-1.0 (EXTERNAL)."""
    lay: Dict[str, tuple] = {}
    if not up_only:
        for a in range(2):
            for t in range(2):
                lay[f"down_blocks.1.attentions.{a}.transformer_blocks.{t}"] = (640, 10)
        for a in range(2):
            for t in range(10):
                lay[f"down_blocks.2.attentions.{a}.transformer_blocks.{t}"] = (1280, 20)
        for t in range(10):
            lay[f"mid_block.attentions.0.transformer_blocks.{t}"] = (1280, 20)
    for a in range(3):
        for t in range(10):
            lay[f"up_blocks.0.attentions.{a}.transformer_blocks.{t}"] = (1280, 20)
    for a in range(3):
        for t in range(2):
            lay[f"up_blocks.1.attentions.{a}.transformer_blocks.{t}"] = (640, 10)
    return lay
