"""Sampling masks of Consistent Self-Attention in compact form.

The reference builds a dense ``(T*N, T*N)`` boolean mask per resolution every denoise step
(``cal_attn_mask_xl``, StoryDiffusion/utils/gradio_utils.py:241-287; 419 MB at 1024², quadratic in frames) although
it is fully determined by ONE sampled vector ``r`` of ``T*N`` booleans: row block ``i`` is
``(r & col < F*N) | col in block_i``.  ``CompactMask`` keeps ``r`` only and turns it into per-frame ascending key
index lists with the ``csa_compact_rows`` CUDA kernel (bit-exact w.r.t. ``torch.nonzero(mask[i*N])``).

RNG parity: ``cal_attn_mask_xl`` below consumes the default generator of ``device`` exactly like the reference —
two ``torch.rand`` calls of shapes ``(1, T*n32)`` then ``(1, T*n16)`` in ``dtype`` (gradio_utils.py:257-258) — so a
pipeline seeded like the reference (Comic_Generation.py:35-40) sees the same sample vectors.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import native


class CompactMask:
    """One resolution's sampling mask: the T distinct rows as index lists, built lazily on the device.

    ``lists()`` returns ``(idx [T, stride] int32, counts [T] int32)``: row ``f < F`` is the key list of write-mode
    frame ``f`` (columns < F*N), row ``F`` is the read-mode list (sampled bank rows followed by the own block
    ``[F*N, T*N)``, whose N entries the attention kernel drops via ``g_adjust = -N``).
    """

    def __init__(self, total_length: int, id_length: int, n_tokens: int, sample: Optional[torch.Tensor] = None,
                 dense: Optional[torch.Tensor] = None):
        if (sample is None) == (dense is None):
            raise ValueError("CompactMask needs exactly one of `sample` (T*N bool vector) or `dense` ((T*N)^2 mask)")
        self.total_length = int(total_length)
        self.id_length = int(id_length)
        self.n_tokens = int(n_tokens)
        self._sample = sample
        self._dense = dense
        self._lists: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._sample_list = None      # (s_idx, s_count, ranges) — see sample_list()
        self._rand = None             # resample_(): the uniform draws the sample vector is thresholded from
        self._sample_pos = None       # sample_positions(): inverse of the sampled list (fused gather of csa_gemm)
        self.shared_sample = sample is not None   # rows are (S u own block) by construction; dense: checked

    # what the reference's mask tensor exposes and drivers may look at
    @property
    def shape(self):
        s = self.total_length * self.n_tokens
        return torch.Size((s, s))

    @property
    def device(self):
        return (self._sample if self._sample is not None else self._dense).device

    def lists(self, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Index lists on ``device`` (default: where the sample lives).  A sample drawn on another device — e.g.
        ``device="cpu"`` processors, whose torch.rand calls consume the CPU generator — is moved once."""
        if self._lists is None:
            T, F, N = self.total_length, self.id_length, self.n_tokens
            if device is not None:
                if self._sample is not None and self._sample.device != torch.device(device):
                    self._sample = self._sample.to(device)
                if self._dense is not None and self._dense.device != torch.device(device):
                    self._dense = self._dense.to(device)
            if self._sample is not None:
                r = self._sample
                if r.dtype != torch.bool or r.numel() != T * N or not r.is_contiguous():
                    raise ValueError("sample vector must be a contiguous bool tensor of T*N elements")
                self._lists = native.compact_rows(r, T, T * N, 0, block_n=N, limit_cols=F * N)
            else:
                m = self._dense
                # the T distinct rows are rows 0, N, 2N, ... of the dense mask (gradio_utils.py:285-286)
                self._lists = native.compact_rows(m, T, T * N, N * m.stride(0))
        return self._lists

    def sample_list(self, device=None):
        """``(s_idx [stride] int32, s_count [1] int32, ranges [T, 4] int32)`` on the device: the ascending list S of
        sampled key positions below ``F*N`` — ONE list shared by all frames (gradio_utils.py:257-261) — and, per
        frame, the two runs of S it attends besides its own block (``csa_sample_ranges``).  Only valid when the mask
        has that structure (``shared_sample``); a dense mask from outside is checked by ``from_dense``."""
        if not self.shared_sample:
            raise ValueError("mask rows do not share one sample vector; use lists()")
        if self._sample_list is None:
            T, F, N = self.total_length, self.id_length, self.n_tokens
            if self._sample is not None:
                src = self._sample
                if device is not None and src.device != torch.device(device):
                    src = self._sample = src.to(device)
                if src.dtype != torch.bool or src.numel() != T * N or not src.is_contiguous():
                    raise ValueError("sample vector must be a contiguous bool tensor of T*N elements")
            else:
                m = self._dense
                if device is not None and m.device != torch.device(device):
                    m = self._dense = m.to(device)
                src = m[F * N]          # the read-mode row: S on columns < F*N (gradio_utils.py:267-278)
            s_idx, s_count = native.compact_rows(src, 1, F * N, 0)
            ranges = native.sample_ranges(s_idx, s_count, N, F)
            self._sample_list = (s_idx.view(-1), s_count, ranges)
        return self._sample_list

    def sample_positions(self, device=None) -> torch.Tensor:
        """``pos [F*N] int32``: position of every key column < F*N in the sampled list S, or -1 — what the K|V
        projection's epilogue (``csa_gemm``) looks a row up in to store it straight into K[S] / V[S]."""
        if self._sample_pos is None:
            s_idx, s_count, _ = self.sample_list(device)
            self._sample_pos = native.sample_positions(s_idx, s_count, self.id_length * self.n_tokens)
        return self._sample_pos

    def resample_(self, sa: float, dtype=torch.float16, post_sample=None) -> "CompactMask":
        """Draw a fresh sample vector INTO the buffers this mask already owns and refresh, in place, whichever index
        lists were built from the old one — what ``cal_attn_mask_xl`` does at every step roll-over
        (Comic_Generation.py:119-125) without a single allocation.  Every device pointer a previous step handed to a
        kernel stays valid and now holds the new step's data, which is what lets a whole denoise step be captured in
        ONE CUDA graph and replayed (``spider_b200.graph.StepGraph``).  Consumes the device's generator exactly like
        ``torch.rand((1, T*n), device, dtype) < sa`` (gradio_utils.py:257-258).  ``post_sample`` (multi-GPU: broadcast
        from rank 0) is applied to the sample vector before the lists are rebuilt."""
        if self._sample is None:
            raise ValueError("resample_() needs a mask in sample-vector form")
        T, F, N = self.total_length, self.id_length, self.n_tokens
        r = self._rand
        if r is None or r.dtype != dtype or r.device != self._sample.device:
            r = self._rand = torch.empty((1, T * N), device=self._sample.device, dtype=dtype)
        torch.rand((1, T * N), device=r.device, dtype=dtype, out=r)
        torch.lt(r[0], sa, out=self._sample)
        if post_sample is not None:
            got = post_sample(self._sample)
            if got is not None and got.data_ptr() != self._sample.data_ptr():
                self._sample.copy_(got)
        if self._lists is not None:
            idx, counts = self._lists
            native.compact_rows(self._sample, T, T * N, 0, block_n=N, limit_cols=F * N, idx=idx, counts=counts)
        if self._sample_list is not None:
            s_idx, s_count, ranges = self._sample_list
            native.compact_rows(self._sample, 1, F * N, 0, idx=s_idx.view(1, -1), counts=s_count)
            native.sample_ranges(s_idx, s_count, N, F, out=ranges)
            if self._sample_pos is not None:
                native.sample_positions(s_idx, s_count, F * N, out=self._sample_pos)
        self._shard_plan = None   # per-rank run lengths read back from the old sample are stale
        return self

    def dense(self) -> torch.Tensor:
        """Materialise the reference's dense mask (debugging / interoperability only; O((T*N)^2) bytes)."""
        if self._dense is not None:
            return self._dense
        T, F, N = self.total_length, self.id_length, self.n_tokens
        rows = self._sample.unsqueeze(0).repeat(T, 1)
        rows[:, F * N:] = False
        for i in range(T):
            rows[i, i * N:(i + 1) * N] = True
        return rows.unsqueeze(1).repeat(1, N, 1).reshape(-1, T * N)


def cal_attn_mask_xl(total_length, id_length, sa32, sa64, height, width, device="cuda", dtype=torch.float16,
                     reuse=None, post_sample=None):
    """Drop-in for the reference's ``cal_attn_mask_xl`` (same signature, same RNG consumption) that returns two
    ``CompactMask`` objects instead of two dense tensors.

    ``reuse=(mask1024, mask4096)``: when both are ``CompactMask`` objects of this geometry that live on ``device``,
    they are re-sampled IN PLACE (``CompactMask.resample_``) and returned — no allocation, stable device pointers
    (CUDA-graph replay); anything else is ignored and fresh masks are built."""
    n32 = (height // 32) * (width // 32)   # gradio_utils.py:250
    n16 = (height // 16) * (width // 16)   # gradio_utils.py:251
    if reuse is not None and _reusable(reuse[0], total_length, id_length, n32, device) and \
            _reusable(reuse[1], total_length, id_length, n16, device):
        reuse[0].resample_(sa32, dtype, post_sample)   # :257
        reuse[1].resample_(sa64, dtype, post_sample)   # :258
        return reuse[0], reuse[1]
    r32 = torch.rand((1, total_length * n32), device=device, dtype=dtype) < sa32   # :257
    r16 = torch.rand((1, total_length * n16), device=device, dtype=dtype) < sa64   # :258
    r32, r16 = r32[0], r16[0]
    if post_sample is not None:
        r32 = post_sample(r32)
        r16 = post_sample(r16)
    return (CompactMask(total_length, id_length, n32, sample=r32),
            CompactMask(total_length, id_length, n16, sample=r16))


def _reusable(cm, total_length, id_length, n_tokens, device) -> bool:
    if not isinstance(cm, CompactMask) or cm._sample is None:
        return False
    dev = torch.device(device)
    sd = cm._sample.device
    same_dev = sd.type == dev.type and (dev.index is None or sd.index == dev.index)
    return (cm.total_length == total_length and cm.id_length == id_length and cm.n_tokens == n_tokens and same_dev
            and cm._sample.dtype == torch.bool and cm._sample.is_contiguous())


def from_dense(mask: torch.Tensor, total_length: int, id_length: int, validate: bool = True) -> CompactMask:
    """Wrap a dense mask produced by the unmodified reference sampler (Comic_Generation.py:376).

    With ``validate`` the premise of the compaction — all N rows of a frame block are identical — is checked on the
    device (HBM-bound pass over the mask, one host sync); a mask that violates it raises ``ValueError`` instead of
    being silently mis-compacted.
    """
    if mask.dim() != 2 or mask.shape[0] != mask.shape[1] or mask.dtype != torch.bool:
        raise ValueError(f"expected a square 2-D bool mask, got {tuple(mask.shape)} {mask.dtype}")
    if mask.shape[0] % total_length:
        raise ValueError(f"mask side {mask.shape[0]} is not a multiple of total_length {total_length}")
    if mask.stride(1) != 1:
        mask = mask.contiguous()
    n = mask.shape[0] // total_length
    if validate:
        bad = int(native.validate_mask(mask, n).item())
        if bad:
            raise ValueError(
                f"attention mask has {bad} 16-byte words that differ between rows of one frame block; consistent "
                "self-attention kernels need the per-frame mask structure of cal_attn_mask_xl")
    cm = CompactMask(total_length, id_length, n, dense=mask)
    if validate:
        # do all rows share one sample vector (row i = S u block_i)?  Only then may the sampled K/V rows be gathered
        # once per layer; any other per-frame mask takes the generic in-kernel gather over per-frame lists.
        rows = mask[::n]                                   # (T, T*n): the T distinct rows
        fn = id_length * n
        want = rows[id_length].clone()
        want[fn:] = False
        want = want.unsqueeze(0).repeat(total_length, 1)
        for i in range(total_length):
            want[i, i * n:(i + 1) * n] = True
        cm.shared_sample = bool(torch.equal(rows, want))
    else:
        cm.shared_sample = True
    return cm
