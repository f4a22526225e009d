"""Low-VRAM flavour of the Consistent Self-Attention processor (the README-recommended StoryDiffusion entry point).

Mirrors ``SpatialAttnProcessor2_0`` of StoryDiffusion/gradio_app_sdxl_specific_id_low_vram.py:99-366 and the sampler
``cal_attn_indice_xl_effcient_memory`` of StoryDiffusion/utils/gradio_utils.py:303-312: same constructor and
``__call__`` signature, same control globals of the host module (``write, cur_step, attn_count, total_count, sa32,
sa64, height, width, indices1024, indices4096, cur_character``; :145-149), same branch logic (early cutoff
``cur_step < 1``, gate 0.3 -> 0.1 at step 20; :192-201), same consumption of Python's ``random`` (one draw per call
iff ``cur_step >= 1``) and of torch's generator (two ``torch.rand`` of shape ``(T, n)`` per step), same id_bank layout
``id_bank[character][step] = [one (2, K_img, C) tensor of sampled hidden tokens per reference image]`` (:172-179), so
that ``save_single_character_weights`` / ``load_single_character_weights`` (:437-479) and code that pokes
``id_bank / id_length / total_length`` (:746-752) keep working.

What changes is how the attention is computed — with the same kernels as the main variant:
  * the reference attends, image by image, ``cat(sampled tokens of the other images, all own tokens)`` (:231-246) —
    i.e. key set ``S' \\ block_f  u  block_f`` with ``S'`` the concatenation of the per-image sampled positions.  That
    is the key-list structure of the main variant with the ``(T, n)`` Bernoulli matrix read as one ``T*n`` sample
    vector, so the write pass is ONE ``csa_gather_kv`` + ONE ``csa_attn_fwd`` launch for all images, heads and CFG
    halves instead of ``img_nums`` SDPA calls over concatenated copies;
  * the bank rows are gathered on the device (``csa_gather_rows`` over the compacted list); the per-image split of the
    packed buffer is only materialised when somebody looks at the list (persistence, debugging);
  * a read frame attends ``bank tokens of its characters + itself`` (:186-190, :252-261) as K/V source A (projected
    bank rows, contiguous) + source B (the current frame), without ``torch.cat`` of the current frame.
There is no CPU / eager fallback.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F_torch

from . import masks as _masks
from . import native
from .processor import SpatialAttnProcessor2_0 as _MainProcessor


class CompactIndices:
    """One resolution's per-frame sampled token positions (what the reference keeps as a list of T index tensors).

    Holds the ``(T, n)`` Bernoulli matrix; the attention path uses it in compact form (``mask_for``), the list
    protocol (``len``, ``[i]``, iteration) gives the reference's index tensors on demand (one ``nonzero`` each)."""

    def __init__(self, matrix: torch.Tensor):
        if matrix.dim() != 2 or matrix.dtype != torch.bool:
            raise ValueError("expected a (T, n) bool matrix")
        self.matrix = matrix.contiguous()
        self.total_length, self.n_tokens = matrix.shape
        self._masks: Dict[int, _masks.CompactMask] = {}

    def mask_for(self, img_nums: int) -> _masks.CompactMask:
        """Key lists of a write batch of ``img_nums`` images: image f attends the sampled tokens of images
        ``g != f, g < img_nums`` plus its own block — ``CompactMask`` with ``id_length = img_nums``."""
        if not 0 < img_nums <= self.total_length:
            raise ValueError(f"{img_nums} images but index lists were sampled for {self.total_length} frames")
        cm = self._masks.get(img_nums)
        if cm is None:
            cm = _masks.CompactMask(self.total_length, img_nums, self.n_tokens, sample=self.matrix.view(-1))
            self._masks[img_nums] = cm
        return cm

    def __len__(self):
        return self.total_length

    def __getitem__(self, i):
        return torch.nonzero(self.matrix[i], as_tuple=True)[0]

    def __iter__(self):
        return (self[i] for i in range(self.total_length))


def cal_attn_indice_xl_effcient_memory(total_length, id_length, sa32, sa64, height, width, device="cuda",
                                       dtype=torch.float16):
    """Drop-in for the reference sampler (gradio_utils.py:303-312): same signature, same two ``torch.rand`` calls,
    returns two ``CompactIndices`` instead of two lists of index tensors (no host sync)."""
    nums_1024 = (height // 32) * (width // 32)
    nums_4096 = (height // 16) * (width // 16)
    bool_matrix1024 = torch.rand((total_length, nums_1024), device=device, dtype=dtype) < sa32
    bool_matrix4096 = torch.rand((total_length, nums_4096), device=device, dtype=dtype) < sa64
    return CompactIndices(bool_matrix1024), CompactIndices(bool_matrix4096)


class SampledBankEntry(list):
    """``id_bank[character][step]``: the sampled hidden tokens of every reference image.

    Kept packed — ``packed (2, cap, C)`` with ``count`` (device int32) valid rows per CFG half and the per-image
    boundaries derivable from ``ranges`` — and presented as the reference's list of ``(2, K_img, C)`` tensors when
    iterated / indexed (one host sync, then cached).  Lists assigned from outside (a loaded checkpoint) are wrapped
    with ``from_list``."""

    def __init__(self, packed=None, count=None, img_nums=0, n_tokens=0, s_idx=None):
        super().__init__()
        self.packed, self.count, self.img_nums, self.n_tokens, self._s_idx = packed, count, img_nums, n_tokens, s_idx
        self._materialised = packed is None
        self.k = self.v = None        # projected rows (filled by the write pass in "kv"/"both" mode, or lazily)
        self._n_valid: Optional[int] = None

    @classmethod
    def from_list(cls, tensors: List[torch.Tensor]):
        e = cls()
        list.extend(e, tensors)
        return e

    def _materialise(self):
        if self._materialised:
            return
        n = self._n_valid if self._n_valid is not None else int(self.count.item())
        pos = self._s_idx[:n].to(torch.int64)
        img = torch.div(pos, self.n_tokens, rounding_mode="floor")
        bounds = torch.searchsorted(img, torch.arange(self.img_nums + 1, device=img.device)).tolist()
        list.extend(self, [self.packed[:, bounds[i]:bounds[i + 1]] for i in range(self.img_nums)])
        self._n_valid = n
        self._materialised = True

    def __len__(self):
        self._materialise()
        return list.__len__(self)

    def __iter__(self):
        self._materialise()
        return list.__iter__(self)

    def __getitem__(self, i):
        self._materialise()
        return list.__getitem__(self, i)

    def hidden_rows(self, device):
        """(2, K, C) valid sampled hidden tokens (all images), for projection."""
        if self.packed is not None:
            if self._n_valid is None:
                self._n_valid = int(self.count.item())
            return self.packed[:, :self._n_valid].to(device)
        return torch.cat([t.to(device) for t in list.__iter__(self)], dim=1)


class SpatialAttnProcessorLowVram(torch.nn.Module):
    r"""Low-VRAM Consistent Self-Attention processor (B200-native).  See the module docstring.

    Class attributes (set on the class or on a bound subclass from ``make_lowvram_processor_class``):
      _host        namespace holding the control globals
      bank_store   "hidden" (default: the reference's layout and memory; K/V of the bank rows are projected on every
                   read call, as the reference does) | "kv" (the write pass also keeps the projected sampled rows it
                   computes anyway: reads need no projection, 3x the bank memory)
    """

    _host = None
    bank_store = "hidden"

    def __init__(self, hidden_size=None, cross_attention_dim=None, id_length=4, device="cuda", dtype=torch.float16):
        super().__init__()
        if not hasattr(F_torch, "scaled_dot_product_attention"):   # :121-124
            raise ImportError("AttnProcessor2_0 requires PyTorch 2.0, to use it, please upgrade PyTorch to 2.0.")
        self.device = device
        self.dtype = dtype
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.total_length = id_length + 1
        self.id_length = id_length
        self.id_bank: Dict[str, Dict[int, list]] = {}

    # ------------------------------------------------------------------------------------------------ helpers
    def _resample(self, h):
        h.indices1024, h.indices4096 = cal_attn_indice_xl_effcient_memory(
            self.total_length, self.id_length, h.sa32, h.sa64, h.height, h.width, device=self.device,
            dtype=self.dtype)

    def _indices(self, h, n_tokens: int, device) -> CompactIndices:
        use32 = n_tokens == (h.height // 32) * (h.width // 32)      # :163-166 / :204-207
        ind = h.indices1024 if use32 else h.indices4096
        if not isinstance(ind, CompactIndices):
            # index lists from the unmodified reference sampler: rebuild the bool matrix once per list object
            cache = getattr(h, "_csa_indices_cache", None)
            if cache is None:
                cache = {}
                setattr(h, "_csa_indices_cache", cache)
            hit = cache.get(use32)
            if hit is None or hit[0] is not ind:
                m = torch.zeros((len(ind), n_tokens), dtype=torch.bool, device=device)
                for i, ix in enumerate(ind):
                    m[i, ix.to(device)] = True
                hit = (ind, CompactIndices(m))
                cache[use32] = hit
            ind = hit[1]
        if ind.n_tokens != n_tokens:
            raise ValueError(f"index lists hold {ind.n_tokens} tokens per frame but hidden_states has {n_tokens} "
                             "(height/width globals do not match the latent size)")
        if ind.matrix.device != torch.device(device):
            ind.matrix = ind.matrix.to(device)
            ind._masks.clear()
        return ind

    def _entry(self, e) -> SampledBankEntry:
        return e if isinstance(e, SampledBankEntry) else SampledBankEntry.from_list(list(e))

    # ------------------------------------------------------------------------------------------------ __call__
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        h = self._host
        if attention_mask is not None or encoder_hidden_states is not None:
            raise NotImplementedError("SpatialAttnProcessor2_0 low-VRAM (B200): self-attention only — "
                                      "attention_mask and encoder_hidden_states must be None")
        _MainProcessor._check_input(hidden_states)
        if hidden_states.ndim != 3:
            raise ValueError(f"hidden_states must be (B, N, C), got {tuple(hidden_states.shape)}")
        if attn.spatial_norm is not None or attn.group_norm is not None:
            raise NotImplementedError("low-VRAM processor: spatial_norm / group_norm are not used by SDXL attn1")
        if h.attn_count == 0 and h.cur_step == 0:                   # :150-160
            self._resample(h)
        B, N, C = hidden_states.shape
        heads = attn.heads
        if C != heads * native.CSA_HEAD_DIM:
            raise native.CsaNativeError(f"head_dim {C // heads} != 64: not an SDXL self-attention layer")
        x = hidden_states.contiguous()
        x2 = x.view(B * N, C)
        write = bool(h.write)
        cur_step = h.cur_step
        chars = list(h.cur_character)
        dev = x.device

        q = attn.to_q(x).view(B * N, C)
        k = attn.to_k(x).view(B * N, C)
        v = attn.to_v(x).view(B * N, C)
        o = torch.empty_like(q)

        img_nums = B // 2
        cm = None
        if write:                                                   # :161-180
            if len(chars) != 1:
                raise AssertionError("write pass expects exactly one current character")
            cm = self._indices(h, N, dev).mask_for(img_nums)
            s_idx, s_count, _ = cm.sample_list(dev)
            cap = img_nums * N
            # the bank keeps the SAMPLED tokens only, like the reference (:174-179, `.clone()` of the gathered rows):
            # exactly `count` rows per CFG half.  The count is read back once per re-sampled mask and shared by all
            # layers of the step (one host sync per step and resolution).
            n_valid = getattr(cm, "_n_sampled_host", None)
            if n_valid is None:
                n_valid = cm._n_sampled_host = int(s_count.item())
            packed = torch.empty((2, n_valid, C), dtype=x.dtype, device=dev)
            if n_valid > 0:
                for g in range(2):
                    native.gather_rows(x2, s_idx, n_valid, row_base=g * cap, count=s_count, out=packed[g])
            entry = SampledBankEntry(packed, s_count, img_nums, N, s_idx)
            entry._n_valid = n_valid
            self.id_bank.setdefault(chars[0], {})[cur_step] = entry
        else:
            if B != 2:
                raise ValueError(f"read pass expects batch 2 (uncond, cond), got {B}")
            entries = [self._entry(self.id_bank[c][cur_step]) for c in chars]   # KeyError like the reference (:189)

        branch = "early"
        if cur_step < 1:                                            # :192-195
            self._standard(q, k, v, o, B, N, heads)
        else:
            random_number = random.random()                         # :197
            rand_num = 0.3 if cur_step < 20 else 0.1                # :198-201
            if random_number > rand_num:
                branch = "consistent"
                if write:
                    s_idx, s_count, ranges = cm.sample_list(dev)
                    k_s, v_s, capg = native.gather_kv(k, v, img_nums * N, 2, s_idx, s_count, img_nums * N)
                    if self.bank_store == "kv":
                        entry.k, entry.v, entry.kv_group_rows = k_s, v_s, capg
                    native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=img_nums, n_q=N,
                                    k_a=k_s, v_a=v_s, a_group_rows=capg, ranges=ranges, range_base=0, range_step=1,
                                    k_b=k, v_b=v, b_group_rows=img_nums * N, cb=(0, N, N))
                else:
                    self._read(attn, entries, q, k, v, o, N, heads, dev)
            else:
                branch = "standard"
                self._standard(q, k, v, o, B, N, heads)             # :263-266
        self._last_branch = branch

        out = attn.to_out[1](attn.to_out[0](o.view(B, N, C)))       # :347-349
        if attn.residual_connection:
            out = out + hidden_states
        if attn.rescale_output_factor != 1.0:
            out = out / attn.rescale_output_factor

        h.attn_count += 1                                           # :267-280
        if h.attn_count == h.total_count:
            h.attn_count = 0
            h.cur_step += 1
            self._resample(h)
        return out

    # ------------------------------------------------------------------------------------------------ branches
    def _standard(self, q, k, v, o, B, N, heads):
        """``__call2__`` with encoder_hidden_states=None (:284-366): plain per-image self-attention."""
        native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=B, n_q=N, k_b=k, v_b=v, b_group_rows=B * N,
                        cb=(0, N, N))

    def _read(self, attn, entries, q, k, v, o, N, heads, dev):
        """Read mode (:247-262): keys = sampled bank tokens of every current character + the frame itself."""
        ks, vs = [], []
        for e in entries:
            if e.k is not None and e.v is not None and e.packed is not None:
                n = e._n_valid if e._n_valid is not None else int(e.count.item())
                e._n_valid = n
                g = e.kv_group_rows
                ks.append(e.k.view(2, g, -1)[:, :n])
                vs.append(e.v.view(2, g, -1)[:, :n])
            else:
                hid = e.hidden_rows(dev).to(q.dtype)                # (2, K, C), :188 `.to(self.device)`
                ks.append(attn.to_k(hid))
                vs.append(attn.to_v(hid))
        kb = ks[0] if len(ks) == 1 else torch.cat(ks, dim=1)
        vb = vs[0] if len(vs) == 1 else torch.cat(vs, dim=1)
        K = kb.shape[1]
        if K == 0:
            return self._standard(q, k, v, o, 2, N, heads)
        kb = kb.contiguous().view(2 * K, -1)
        vb = vb.contiguous().view(2 * K, -1)
        native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=1, n_q=N,
                        k_a=kb, v_a=vb, a_group_rows=K, ca=(0, 0, K),
                        k_b=k, v_b=v, b_group_rows=N, cb=(0, 0, N))


_LOWVRAM_DEFAULTS = dict(write=False, cur_step=0, attn_count=0, total_count=0, sa32=0.5, sa64=0.5, height=768,
                         width=768, indices1024=None, indices4096=None, cur_character=[])


def make_lowvram_processor_class(host, bank_store: Optional[str] = None):
    ns = {"_host": host, "__doc__": SpatialAttnProcessorLowVram.__doc__,
          "__module__": SpatialAttnProcessorLowVram.__module__}
    if bank_store is not None:
        if bank_store not in ("hidden", "kv"):
            raise ValueError("bank_store must be 'hidden' or 'kv'")
        ns["bank_store"] = bank_store
    return type("SpatialAttnProcessor2_0", (SpatialAttnProcessorLowVram,), ns)


def install_lowvram(host, replace_sampler: bool = True, bank_store: Optional[str] = None):
    """Rebind ``host.SpatialAttnProcessor2_0`` (and ``host.cal_attn_indice_xl_effcient_memory``) of the low-VRAM
    application module (gradio_app_sdxl_specific_id_low_vram.py instantiates the class by name at :388, :605)."""
    for name, val in _LOWVRAM_DEFAULTS.items():
        if not hasattr(host, name):
            setattr(host, name, list(val) if isinstance(val, list) else val)
    cls = make_lowvram_processor_class(host, bank_store)
    if hasattr(host, "SpatialAttnProcessor2_0") and not hasattr(host, "_csa_original_processor"):
        host._csa_original_processor = host.SpatialAttnProcessor2_0
    host.SpatialAttnProcessor2_0 = cls
    if replace_sampler:
        if hasattr(host, "cal_attn_indice_xl_effcient_memory") and not hasattr(host, "_csa_original_index_sampler"):
            host._csa_original_index_sampler = host.cal_attn_indice_xl_effcient_memory
        host.cal_attn_indice_xl_effcient_memory = cal_attn_indice_xl_effcient_memory
    return cls


# ---------------------------------------------------------------------------------------------------- persistence
def save_single_character_weights(unet, character, description, filepath):
    """Same on-disk format as the reference (:437-457): ``{description, character, <processor name>: {step: [cpu
    tensors]}}`` with one ``(2, K_img, C)`` tensor of sampled hidden tokens per reference image — files are
    interchangeable with the reference's (``tests/golden/lowvram_bank_bob.pt`` was written by the reference's own
    function).  Needs ``bank_store="hidden"`` entries (the default) or loaded lists."""
    blob = {"description": description, "character": character}
    for name, proc in unet.attn_processors.items():
        if not isinstance(proc, SpatialAttnProcessorLowVram):
            continue
        blob[name] = {step: [t.cpu().clone() for t in entry] for step, entry in proc.id_bank[character].items()}
    torch.save(blob, filepath)


def load_single_character_weights(unet, filepath):
    """Mirror of the reference loader (:460-479): fills ``id_bank[character]`` of every low-VRAM processor."""
    blob = torch.load(filepath, map_location=torch.device("cpu"))
    character = blob["character"]
    device = getattr(unet, "device", None)
    for name, proc in unet.attn_processors.items():
        if not isinstance(proc, SpatialAttnProcessorLowVram):
            continue
        proc.id_bank[character] = {
            step: SampledBankEntry.from_list([t.to(device) if device is not None else t for t in tensors])
            for step, tensors in blob[name].items()}
    return character, blob["description"]
