"""The reference's stock SDPA processor (StoryDiffusion/utils/gradio_utils.py:387-472).

CPU: the oracle restatement (oracle.reference_port.stock_attention) against tests/golden/stock.npz, outputs of the
UNMODIFIED reference class (tests/golden/make_golden_stock.py).  GPU (`-m gpu`): spider_b200.AttnProcessor2_0 — the
contiguous-segment mode of csa_attn_fwd + csa_gemm projections — against the same fixtures, fp16 and bf16, self- and
cross-attention (77 text tokens: a ragged key tile), 4-D input, residual / rescale; and its installation under the
name the reference instantiates it by."""
import os
import types

import numpy as np
import pytest
import torch

import spider_b200
from spider_b200 import native
from oracle import reference_port as rp
from oracle.fake_diffusers import FakeAttention, FakeUNet

from helpers import MAX_ABS, MIN_COS, max_abs_cos

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stock.npz"))
N_CASES = int(Z["n_cases"])


def _case(ci, dtype=torch.float32, device="cpu"):
    w = {k[len(f"c{ci}_w_"):]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(f"c{ci}_w_")}
    C = w["to_q.weight"].shape[0]
    x = torch.from_numpy(Z[f"c{ci}_x"])
    y = torch.from_numpy(Z[f"c{ci}_y"])
    enc = torch.from_numpy(Z[f"c{ci}_enc"]) if f"c{ci}_enc" in Z.files else None
    heads = C // 64
    attn = FakeAttention(C, heads)
    if enc is not None:
        attn.to_k = torch.nn.Linear(enc.shape[-1], C, bias=False)
        attn.to_v = torch.nn.Linear(enc.shape[-1], C, bias=False)
    attn.load_state_dict(w)
    name = str(Z[f"c{ci}_name"])
    attn.residual_connection = name in ("self4d", "cross_ragged")
    attn.rescale_output_factor = 2.0 if name == "self4d" else 1.0
    attn = attn.to(device=device, dtype=dtype)
    return name, attn, x, enc, y


@pytest.mark.parametrize("ci", range(N_CASES))
def test_oracle_stock_attention_matches_the_reference_class(ci):
    name, attn, x, enc, want = _case(ci)
    with torch.no_grad():
        got = rp.stock_attention(attn, x, enc)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-5, name


def test_install_rebinds_the_stock_processor_name():
    host = types.SimpleNamespace(AttnProcessor=object)
    spider_b200.install(host)
    assert host.AttnProcessor is spider_b200.AttnProcessor2_0 and host._csa_original_stock_processor is object
    p = host.AttnProcessor()                       # the reference calls it without arguments (:368)
    assert isinstance(p, torch.nn.Module)
    spider_b200.AttnProcessor2_0(hidden_size=1280, cross_attention_dim=2048)   # and diffusers' keyword form
    spider_b200.uninstall(host)
    assert host.AttnProcessor is object
    unet = FakeUNet({"down_blocks.1.attentions.0.transformer_blocks.0": (128, 2),
                     "up_blocks.0.attentions.0.transformer_blocks.0": (128, 2)}, with_cross=True)
    n = spider_b200.set_attention_processor(unet, id_length=4, other_processor="b200")
    procs = unet.attn_processors
    assert n == 1
    assert sum(isinstance(p, spider_b200.AttnProcessor2_0) for p in procs.values()) == len(procs) - 1
    with pytest.raises(native.CsaNativeError):     # no CPU fallback
        spider_b200.AttnProcessor2_0()(FakeAttention(128, 2), torch.zeros(1, 16, 128))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("ci", range(N_CASES))
def test_b200_stock_processor_matches_the_reference_class(ci, dtype):
    name, attn, x, enc, want = _case(ci, dtype=dtype, device="cuda:0")
    proc = spider_b200.AttnProcessor2_0()
    before = dict(native.LAUNCHES)
    with torch.no_grad():
        got = proc(attn, x.to("cuda:0", dtype), encoder_hidden_states=None if enc is None else enc.to("cuda:0", dtype))
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == dtype
    assert native.LAUNCHES["csa_attn_fwd"] - before["csa_attn_fwd"] == 1
    err, cos = max_abs_cos(got, want)
    assert err <= MAX_ABS and cos >= MIN_COS, f"{name} {dtype}: max-abs {err:.3e} cos {cos:.6f}"
    with pytest.raises(NotImplementedError):
        proc(attn, x.to("cuda:0", dtype), attention_mask=torch.zeros(1, device="cuda:0"))


@pytest.mark.gpu
def test_b200_stock_processor_sdxl_cross_attention_shape():
    """The SDXL cross-attention shape: 8 latents x 1024 tokens x 1280 channels against 77 text tokens of width 2048
    (K/V projections on csa_gemm with K = 2048), vs torch SDPA in fp32 on the same bf16 inputs."""
    dev, dtype = "cuda:0", torch.bfloat16
    B, N, C, heads, Nk, D = 8, 1024, 1280, 20, 77, 2048
    g = torch.Generator(device=dev).manual_seed(3)
    attn = FakeAttention(C, heads)
    attn.to_k = torch.nn.Linear(D, C, bias=False)
    attn.to_v = torch.nn.Linear(D, C, bias=False)
    attn = attn.to(dev, dtype)
    x = torch.randn((B, N, C), device=dev, generator=g).to(dtype)
    enc = torch.randn((B, Nk, D), device=dev, generator=g).to(dtype)
    before = dict(native.LAUNCHES)
    with torch.no_grad():
        got = spider_b200.AttnProcessor2_0()(attn, x, encoder_hidden_states=enc)
        want = rp.stock_attention(attn.float(), x.float(), enc.float())
    assert native.LAUNCHES["csa_gemm"] - before["csa_gemm"] == 4          # q, k, v, out: no library GEMM
    err, cos = max_abs_cos(got, want)
    assert err <= MAX_ABS and cos >= MIN_COS, f"max-abs {err:.3e} cos {cos:.6f}"
