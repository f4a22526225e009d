"""Drop-in ``SpatialAttnProcessor2_0`` — the diffusers attention-processor plugin of StoryDiffusion's Consistent
Self-Attention, re-implemented on top of ``libcsa_b200.so`` (hand-written sm_100a kernels).

Mirrors the reference class (StoryDiffusion/Comic_Generation.py:46-268; identical copies in app.py:66-294,
predict.py:135-409): same constructor and ``__call__(attn, hidden_states, encoder_hidden_states, attention_mask,
temb)`` signature, same public attributes (``id_bank, id_length, total_length, device, dtype``), same control
surface — the module globals ``write, cur_step, attn_count, total_count, sa32, sa64, height, width, mask1024,
mask4096`` of the *host* module (:82-85) — same branch logic and the same consumption of Python's ``random`` and of
torch's generator, so a pipeline seeded like the reference walks through the same gates and sample vectors.

What changes is how the attention is computed:
  * the dense ``(T*N)^2`` mask is never used as a mask: its T distinct rows are compacted once per step into key
    index lists (``CompactMask``), and regenerated masks are sampled directly in compact form;
  * ``F.scaled_dot_product_attention(q, k, v, attn_mask=...)`` (:175-177, :248-250) becomes one launch of the
    tcgen05/TMEM flash-attention kernel over TMA-gathered keys (``native.attn_fwd``);
  * the ``torch.cat`` of bank and current frames (:92) and the re-projection of the bank rows (:162-165) disappear:
    the write pass keeps the projected K/V it computes anyway (``bank.py``) and the kernel reads bank and current
    K/V in place as two sources.
The four ``nn.Linear`` projections stay with the ``attn`` module that owns the weights (cuBLAS).

There is no CPU / eager fallback: non-CUDA tensors, non 16-bit dtypes or a missing library raise.
"""
from __future__ import annotations

import random
from typing import Optional

import torch
import torch.nn.functional as F_torch

from . import masks as _masks
from . import native
from .bank import STORE_MODES, BankEntry, IdBank


class BankCapacityError(RuntimeError):
    """The preallocated id_bank arena of a layer is full (``bank_capacity`` steps)."""


class StoryGlobals:
    """Default host namespace: the reference's module globals (Comic_Generation.py:82-85, :327-349) as attributes.
    ``install()`` binds processors to the reference's own module instead."""

    def __init__(self):
        self.write = False
        self.cur_step = 0
        self.attn_count = 0
        self.total_count = 0
        self.sa32 = 0.5
        self.sa64 = 0.5
        self.height = 768
        self.width = 768
        self.id_length = 4
        self.total_length = 5
        self.mask1024 = None
        self.mask4096 = None


GLOBALS = StoryGlobals()


class SpatialAttnProcessor2_0(torch.nn.Module):
    r"""Consistent Self-Attention processor (B200-native).  See the module docstring.

    Class attributes that configure every instance (set them on the class, or on a bound subclass returned by
    ``install.make_processor_class``):
      _host            namespace holding the control globals (a module or any object with those attributes)
      bank_store       "kv" | "hidden" | "both"  — what the write pass keeps per step (``bank.py``)
      validate_masks   check dense masks supplied from outside for the per-frame structure (one sync per mask)
      kv_gather        "pre" | "inline" — how the sampled cross-frame K/V rows reach shared memory
    """

    _host = GLOBALS
    bank_store = "kv"
    validate_masks = True
    kv_gather = "pre"       # "pre": sampled K/V rows gathered once per layer (csa_gather_kv) and streamed as plain
                            # TMA tiles; "inline": per-frame index lists, TMA gather4 inside the attention kernel
    native_gemm = True      # the projections run on the hand-written sm_100a GEMM (csa_gemm) when the shape allows
                            # (N % 128 == 0, K % 64 == 0: every SDXL attn1 layer), else on cuBLASLt (csa_linear)
    gemm_qo = True          # to_q / to_out[0] on the hand-written GEMM too (False: cuBLASLt for those two, a few us
                            # faster per layer at F = 4 — profiles/r02_gemm.md — at the price of library launches)
    fused_qkv = True        # to_q and to_k|to_v as ONE GEMM over the stacked weight [w_q; w_k; w_v] with two outputs
                            # (q, and K|V where the id_bank keeps it): one launch less per layer, fewer partial waves
    fused_gather = True     # consistent write pass: the K|V projection's epilogue stores the sampled rows straight
                            # into K[S] / V[S] (no csa_gather_kv launch, no second pass over K and V)
    native_projections = True   # to_q / to_k|to_v (one GEMM) / to_out[0] issued by the library
                            # (csa_linear) and the whole call handed over in one batch (csa_run_batch) whenever the
                            # attn module is the plain SDXL attn1 shape (bias-free Linear q/k/v, Linear + Dropout(0)
                            # out, no norms); anything else goes through the module's own projections
    bank_capacity = None    # None: the bank grows step by step like the reference's dict (Comic_Generation.py:89), each
                            # entry aliasing that step's K|V projection output.  int: a preallocated arena of that
                            # many steps per layer — the K|V GEMM writes straight into the step's slot, nothing is
                            # allocated during the write pass, and a step beyond the capacity raises BankCapacityError
                            # instead of running the device out of memory (1024^2, id_length 4, "kv": 1.76 GB per
                            # denoise step over the 36 layers, 88 GB for 50 steps; "hidden" halves it)
    inplace_masks = True    # step roll-over re-samples the host's CompactMasks IN PLACE when it can (no allocation,
                            # stable device pointers: a captured denoise step can be replayed, spider_b200/graph.py)
    batched_read = False    # opt-in: a read call may carry R generated frames (batch 2*R, [uncond R,
                            # cond R]); each attends bank + itself exactly like a batch-2 call, in ONE launch.  The
                            # reference generates them one pipe() call at a time (Comic_Generation.py:445-448), so a
                            # batched call consumes one gate draw where R separate calls consume R.

    def __init__(self, hidden_size=None, cross_attention_dim=None, id_length=4, device="cuda", dtype=torch.float16):
        super().__init__()
        if not hasattr(F_torch, "scaled_dot_product_attention"):   # Comic_Generation.py:64-65
            raise ImportError("AttnProcessor2_0 requires PyTorch 2.0, to use it, please upgrade PyTorch to 2.0.")
        self.device = device
        self.dtype = dtype
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.total_length = id_length + 1
        self.id_length = id_length
        self.id_bank = IdBank()
        self.dist = None   # optional frame sharding across GPUs (spider_b200.dist.FrameSharding)

    # ------------------------------------------------------------------------------------------------ helpers
    def _compact_mask(self, h, n_tokens: int) -> _masks.CompactMask:
        """mask1024 / mask4096 selection of Comic_Generation.py:105-114, returned in compact form."""
        use32 = n_tokens == (h.height // 32) * (h.width // 32)
        m = h.mask1024 if use32 else h.mask4096
        if isinstance(m, _masks.CompactMask):
            cm = m
        elif isinstance(m, torch.Tensor):
            # dense mask from the unmodified driver (Comic_Generation.py:376): compact it once, reuse it for every
            # layer of the step (the processors of all layers share the host's mask tensors)
            cache = getattr(h, "_csa_dense_cache", None)
            if cache is None:
                cache = {}
                setattr(h, "_csa_dense_cache", cache)
            key = (m.data_ptr(), m._version, tuple(m.shape))
            hit = cache.get(use32)
            if hit is not None and hit[0] == key:
                cm = hit[1]
            else:
                if not m.is_cuda:   # e.g. a driver that sampled on the CPU generator: move once, compact on GPU
                    m = m.to(self._last_device)
                cm = _masks.from_dense(m, self.total_length, self.id_length, validate=self.validate_masks)
                cache[use32] = (key, cm)
        else:
            raise TypeError(f"mask1024/mask4096 must be a torch.Tensor or CompactMask, got {type(m)}")
        if cm.total_length != self.total_length or cm.id_length != self.id_length:
            raise ValueError(
                f"mask was sampled for total_length={cm.total_length}, id_length={cm.id_length} but the processor "
                f"has total_length={self.total_length}, id_length={self.id_length}")
        if cm.n_tokens != n_tokens:
            raise ValueError(
                f"mask holds {cm.n_tokens} tokens per frame but hidden_states has {n_tokens} "
                "(height/width globals do not match the latent size)")
        return cm

    @staticmethod
    def _check_input(x: torch.Tensor) -> None:
        if not x.is_cuda:
            raise native.CsaNativeError(
                "SpatialAttnProcessor2_0 (B200) got CPU hidden_states: this path has no CPU fallback")
        if x.dtype not in (torch.float16, torch.bfloat16):
            raise native.CsaNativeError(
                f"SpatialAttnProcessor2_0 (B200) computes in fp16/bf16; got {x.dtype} — run the pipeline in half "
                "precision as the reference does (Comic_Generation.py:313)")

    # ------------------------------------------------------------------------------------------------ __call__
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        h = self._host
        if attention_mask is not None:
            raise NotImplementedError("SpatialAttnProcessor2_0 (B200): a caller-supplied attention_mask is not "
                                      "supported (self-attention layers of SDXL never pass one)")
        if encoder_hidden_states is not None:
            raise NotImplementedError("SpatialAttnProcessor2_0 (B200) is a self-attention processor: "
                                      "encoder_hidden_states must be None (the id_bank provides the extra keys)")
        self._check_input(hidden_states)
        self._last_device = hidden_states.device

        residual = hidden_states
        if attn.spatial_norm is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b4, c4, h4, w4 = hidden_states.shape
            hidden_states = hidden_states.view(b4, c4, h4 * w4).transpose(1, 2)
        if hidden_states.ndim != 3:
            raise ValueError(f"hidden_states must be (B, N, C) or (B, C, H, W), got {tuple(hidden_states.shape)}")
        B, N, C = hidden_states.shape
        heads = attn.heads
        if C != heads * native.CSA_HEAD_DIM:
            raise native.CsaNativeError(f"head_dim {C // heads} != 64: not an SDXL self-attention layer")
        x = hidden_states
        if attn.group_norm is not None:
            x = attn.group_norm(x.transpose(1, 2)).transpose(1, 2)
        x = x.contiguous()
        Fl = self.id_length
        write = bool(h.write)
        cur_step = h.cur_step

        # projections (the reference's to_q/to_k/to_v calls, :155,164-165 / :230,237-238); K/V of the current
        # input are needed by every branch
        plan = self._native_plan(attn, x) if self.native_projections else None
        use_native = plan is not None
        if use_native:
            # from here to flush_batch() the library calls are only collected; they are issued in one go
            native.begin_batch(x)
            try:
                x2 = x.view(B * N, C)
                # q, and K and V side by side in one (rows, 2C) buffer — the bank arena's slot of this step when the
                # write pass keeps K/V there.  The projections themselves are issued once the branch is known
                # (_attend_and_project): one stacked-weight GEMM whose epilogue, in the consistent write pass, also
                # fills K[S] / V[S].
                q = torch.empty((B * N, C), dtype=x.dtype, device=x.device)
                kv = self._arena_slot(cur_step, B * N, 2 * C, x) if write and self.bank_capacity else None
                if kv is None:
                    kv = torch.empty((B * N, 2 * C), dtype=x.dtype, device=x.device)
                k, v = kv[:, :C], kv[:, C:]
                kv_job = (x2, plan, q, kv)
            except Exception:
                native.abort_batch()
                raise
        else:
            q = attn.to_q(x).view(B * N, C)
            k = attn.to_k(x).view(B * N, C)
            v = attn.to_v(x).view(B * N, C)
            kv_job = None
        try:
            return self._attend_and_project(attn, h, hidden_states, residual, input_ndim, q, k, v, B, N, C, heads,
                                            write, cur_step, plan, kv_job)
        except Exception:
            native.abort_batch()
            raise

    def _proj(self, x, w, bias=None, out=None, scatter=None, kv=False, exchange=None):
        """One projection: the hand-written GEMM when the shape allows, else cuBLASLt (never with a fused gather)."""
        if self.native_gemm and (kv or self.gemm_qo) and native.gemm_supported(x.shape[0], w.shape[0], x.shape[1]):
            return native.gemm(x, w, bias, out=out, scatter=scatter, exchange=exchange)
        if scatter is not None:
            raise native.CsaNativeError("fused gather needs the hand-written GEMM (shape not supported)")
        return native.linear(x, w, bias, out=out)

    _KS_BUFFERS: dict = {}

    @classmethod
    def _ks_buffers(cls, device, rows: int, C: int, dtype):
        """The (K[S], V[S]) buffer pair of one layer shape on one device: shared by all layers of that shape (they run
        one after the other on the stream), zero once — rows past the current sample count keep finite values of an
        earlier step and are masked by the kernel's key counts."""
        key = (device.index, rows, C, dtype)
        hit = cls._KS_BUFFERS.get(key)
        if hit is None:
            native.flush_batch()   # torch.zeros launches a fill: nothing deferred may be overtaken
            hit = cls._KS_BUFFERS[key] = (torch.zeros((rows, C), dtype=dtype, device=device),
                                          torch.zeros((rows, C), dtype=dtype, device=device))
        return hit

    def _arena_slot(self, step, rows: int, cols: int, like: torch.Tensor):
        """K|V buffer ``(rows, 2C)`` of ``step`` inside this layer's preallocated arena (``bank_capacity`` steps), or
        None when this write pass does not keep K/V.  Allocated once, at the first write call of the layer."""
        if self.bank_store not in ("kv", "both"):
            return None
        cap = int(self.bank_capacity)
        st = self.__dict__.get("_arena")
        if st is None or st[0].shape[1:] != (rows, cols) or st[0].dtype != like.dtype or st[0].device != like.device:
            native.flush_batch()
            st = self.__dict__["_arena"] = (torch.empty((cap, rows, cols), dtype=like.dtype, device=like.device), {})
        buf, slots = st
        i = slots.get(step)
        if i is None:
            if len(slots) >= cap:
                nbytes = buf[0].numel() * buf.element_size()
                raise BankCapacityError(
                    f"id_bank arena of this layer holds {cap} steps x {nbytes / 2 ** 20:.0f} MiB (bank_capacity={cap}); "
                    f"step {step} does not fit — raise bank_capacity, use bank_store='hidden', or clear id_bank")
            i = slots[step] = len(slots)
        return buf[i]

    def clear_bank(self) -> None:
        """Forget every written step (the arena, if any, is kept and reused)."""
        self.id_bank.clear()
        st = self.__dict__.get("_arena")
        if st is not None:
            st[1].clear()

    def invalidate_native_plan(self) -> None:
        """Drop the cached stacked weight ``[w_q; w_k; w_v]``.  Needed only after an edit the cache key cannot see: an
        in-place update THROUGH ``.data`` (``weight.data.copy_(...)``, a manual LoRA merge) changes neither the
        storage pointer nor the version counter of the parameter."""
        self.__dict__.pop("_nplan", None)

    def _native_plan(self, attn, x):
        """``(key, w_q, w_kv, w_out, b_out, w_qkv)`` when the attn module is the plain SDXL attn1 shape — bias-free Linear
        q/k/v, ``[Linear, Dropout]`` output with the dropout inactive, weights of x's dtype on x's device — else
        None (the module's own projections are used).  ``w_kv = cat(to_k.weight, to_v.weight)``.  Cached per attn
        module and re-validated cheaply: the weights' storage and version counters."""
        lin = torch.nn.Linear
        if not x.is_cuda:
            return None
        mods = attn._modules
        tq, tk, tv, out = mods.get("to_q"), mods.get("to_k"), mods.get("to_v"), mods.get("to_out")
        if type(tq) is not lin or type(tk) is not lin or type(tv) is not lin or out is None:
            return None
        om = out._modules
        o0, o1 = om.get("0"), om.get("1")
        if len(om) != 2 or type(o0) is not lin or type(o1) is not torch.nn.Dropout:
            return None
        wq, wk, wv, wo, bo = tq.weight, tk.weight, tv.weight, o0.weight, o0.bias
        key = (id(attn), x.dtype, x.device, wq.data_ptr(), wq._version, wk.data_ptr(), wk._version, wv.data_ptr(),
               wv._version, wo.data_ptr(), wo._version, None if bo is None else bo.data_ptr(),
               o1.p != 0.0 and o1.training)
        hit = self.__dict__.get("_nplan")
        if hit is not None and hit[0] == key:
            return hit if hit[1] is not None else None
        ok = (tq.bias is None and tk.bias is None and tv.bias is None and not key[-1]
              and all(w.dtype == x.dtype and w.device == x.device and w.is_contiguous() for w in (wq, wk, wv, wo))
              and (bo is None or (bo.dtype == x.dtype and bo.is_contiguous())))
        if ok:
            native.flush_batch()     # torch.cat launches a kernel: nothing deferred may be overtaken
            w_qkv = torch.cat([wq.detach(), wk.detach(), wv.detach()], dim=0).contiguous()
            hit = (key, w_qkv[:wq.shape[0]], w_qkv[wq.shape[0]:], wo.detach(), None if bo is None else bo.detach(),
                   w_qkv)
        else:
            hit = (key, None, None, None, None, None)
        self.__dict__["_nplan"] = hit
        return hit if ok else None

    def _attend_and_project(self, attn, h, hidden_states, residual, input_ndim, q, k, v, B, N, C, heads, write,
                            cur_step, plan, kv_job=None):
        Fl = self.id_length
        use_native = plan is not None
        entry: Optional[BankEntry] = None
        if write:
            # :87-89 — keep what the read passes need.  K/V are this call's projection outputs (zero-copy).
            if self.dist is None and B < Fl:
                raise ValueError(f"write pass needs at least id_length={Fl} frames, got batch {B}")
            keep_hidden = self.bank_store in ("hidden", "both")
            keep_kv = self.bank_store in ("kv", "both")
            entry = BankEntry(hidden_states[:Fl] if keep_hidden else None,
                              hidden_states[Fl:] if keep_hidden else None,
                              k if keep_kv else None, v if keep_kv else None)
            self.id_bank[cur_step] = entry
        else:
            entry = self.id_bank[cur_step]   # KeyError for a step the write pass never reached (:92)
            if self.dist is not None and entry.has_kv() and entry.k.shape[0] == self.dist.local_batch * N:
                # the write pass was sharded: this rank holds the K/V of its own frames only.  Make the entry whole
                # (all-gather over the ranks, once per layer and step); reads then need no exchange at all.
                entry.k, entry.v = self.dist.gather_bank(entry.k, entry.v)
            if B != 2 and not (self.batched_read and B > 0 and B % 2 == 0):
                raise ValueError(f"read pass expects batch 2 (uncond, cond), got {B}"
                                 + ("" if self.batched_read else " (set batched_read=True for 2*R frames)"))

        o = torch.empty_like(q)

        def project_kv(scatter=None, exchange=None):
            # the q and K|V projections of the current input, deferred until the branch is known (native path only)
            if kv_job is None:
                return
            x2, pl, q_out, kv_out = kv_job
            if (self.native_gemm and self.fused_qkv and self.gemm_qo and pl[5] is not None
                    and native.gemm_supported(x2.shape[0], 3 * C, C)):
                sc = None if scatter is None else (scatter[0], scatter[1], scatter[2], scatter[3], scatter[4], 2 * C, C)
                native.gemm(x2, pl[5], out=q_out, out2=kv_out, scatter=sc, exchange=exchange)
            else:
                self._proj(x2, pl[1], out=q_out)
                self._proj(x2, pl[2], out=kv_out, scatter=scatter, kv=True, exchange=exchange)

        branch = "early"
        if cur_step < 5:                                        # :94-96
            project_kv()
            if write:
                self._attn_standard(q, k, v, o, B, N, heads)
            else:
                self._attn_read(attn, entry, q, k, v, o, N, heads, cm=None, plan=plan)
        else:
            random_number = random.random()                     # :98 — exactly one draw per call
            rand_num = 0.3 if cur_step < 20 else 0.1            # :99-102
            if random_number > rand_num:
                branch = "consistent"
                cm = self._compact_mask(h, N)
                if write:
                    want_b = 2 * Fl if self.dist is None else self.dist.local_batch
                    if B != want_b:
                        raise ValueError(f"consistent write pass expects batch 2*id_length={2 * Fl}"
                                         f"{'' if self.dist is None else f' sharded to {want_b} per rank'}, got {B} "
                                         "(the reference fails with a mask shape error here)")
                    fused = None
                    local_only = self.dist is None or self.dist.gc == 1   # gc == 1: one whole CFG half per rank
                    if (kv_job is not None and self.fused_gather and self.native_gemm and local_only
                            and self.kv_gather == "pre" and cm.shared_sample
                            and native.gemm_supported(B * N, 2 * C, C)):
                        # the K|V GEMM's epilogue stores the sampled rows into K[S] / V[S] (no gather launch)
                        cap = Fl * N + native.CSA_TILE
                        pos = cm.sample_positions(q.device)
                        k_s, v_s = self._ks_buffers(q.device, 2 * cap, C, q.dtype)
                        fused = (k_s, v_s, cap)
                        project_kv(scatter=(pos, k_s, v_s, Fl * N, cap, C))
                    elif (kv_job is not None and self.fused_gather and self.native_gemm and self.dist is not None
                          and self.dist.can_fuse_exchange(cm) and native.gemm_supported(B * N, 2 * C, C)):
                        # multi-GPU: the same epilogue stores the sampled rows into EVERY peer's K[S] / V[S] over
                        # NVLink and raises this rank's arrival flag there (no scatter launch, no signal launch)
                        fused = self.dist.begin_exchange(cm, N, C, q.dtype, q.device, Fl)
                        project_kv(scatter=(fused["pos"], None, None, B * N, 0, C), exchange=fused["exchange"])
                    else:
                        project_kv()
                    self._attn_write(q, k, v, o, N, heads, cm, fused)
                else:
                    project_kv()
                    self._attn_read(attn, entry, q, k, v, o, N, heads, cm=cm, plan=plan)
            else:
                branch = "standard"
                project_kv()
                self._attn_standard(q, k, v, o, B, N, heads)    # :118 — bank ignored even when reading
        self._last_branch = branch

        if use_native:
            out = self._proj(o, plan[3], plan[4]).view(B, N, C)      # :185 / :256
            native.flush_batch()                                # everything collected since the projections: ONE call
        else:
            out = o.view(B, N, C)
            out = attn.to_out[0](out)                           # :185 / :256
            out = attn.to_out[1](out)
        if input_ndim == 4:
            b4, c4, h4, w4 = residual.shape
            out = out.transpose(-1, -2).reshape(b4, c4, h4, w4)
        if attn.residual_connection:
            out = out + residual
        if attn.rescale_output_factor != 1.0:
            out = out / attn.rescale_output_factor

        h.attn_count += 1                                       # :119-125
        if h.attn_count == h.total_count:
            h.attn_count = 0
            h.cur_step += 1
            native.flush_batch()
            if self.dist is not None:
                self.dist.end_step()
            h.mask1024, h.mask4096 = _masks.cal_attn_mask_xl(
                self.total_length, self.id_length, h.sa32, h.sa64, h.height, h.width,
                device=self.device, dtype=self.dtype,
                reuse=(h.mask1024, h.mask4096) if self.inplace_masks else None,
                # every rank must compact the same sample vector: rank 0's draw, before the lists are rebuilt
                post_sample=self.dist.sync_sample if self.dist is not None else None)
        return out

    # ------------------------------------------------------------------------------------------------ branches
    def _attn_standard(self, q, k, v, o, B, N, heads):
        """__call2__ with encoder_hidden_states=None (:198-268): plain per-frame self-attention."""
        native.attn_fwd(q, o, heads=heads, n_groups=1, n_frames=B, n_q=N,
                        k_b=k, v_b=v, b_group_rows=B * N, cb=(0, N, N))

    def _attn_write(self, q, k, v, o, N, heads, cm, fused=None):
        """__call1__ in write mode (:129-196 with mask[:F*N,:F*N]): frame f attends its own block plus the sampled
        rows of all identity frames of the same CFG half."""
        Fl = self.id_length
        if self.dist is not None:
            return self.dist.attn_write(q, k, v, o, N, heads, cm, Fl, exchanged=fused)
        if self.kv_gather == "pre" and cm.shared_sample:
            # mask row f = S u block_f: gather K[S], V[S] once (HBM-bound), then frame f attends the two runs of that
            # buffer that lie outside its own block + its own block in place
            s_idx, s_count, ranges = cm.sample_list(q.device)
            if fused is not None:
                k_s, v_s, cap = fused         # filled by the K|V projection's epilogue
            else:
                k_s, v_s, cap = native.gather_kv(k, v, Fl * N, 2, s_idx, s_count, Fl * N)
            native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N,
                            k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=ranges, range_base=0, range_step=1,
                            k_b=k, v_b=v, b_group_rows=Fl * N, cb=(0, N, N))
            return
        idx, counts = cm.lists(q.device)
        native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=Fl, n_q=N,
                        k_a=k, v_a=v, a_group_rows=Fl * N,
                        idx=idx, counts=counts, list_base=0, list_step=1)

    def _project_sampled_bank(self, attn, entry, plan, cm, N, q):
        """``bank_store="hidden"`` read, gather-before-project: the consistent read attends only the SAMPLED bank rows
        (mask[F*N:], Comic_Generation.py:106-108), so only those — about sa32/sa64 of the 2*F*N rows the reference
        re-projects on every call (:162-165) — go through to_k|to_v: gather the sampled hidden rows of both CFG
        halves (csa_gather_rows), one K|V GEMM over them, and the result already IS the K[S], V[S] buffer pair the
        attention kernel streams (zero tail included: bias-free projections map the zero rows to zero)."""
        Fl = self.id_length
        C = q.shape[1]
        dev = q.device
        s_idx, s_count, ranges = cm.sample_list(dev)
        cap = Fl * N + native.CSA_TILE
        g = torch.zeros((2 * cap, C), dtype=q.dtype, device=dev)
        for half, hs in enumerate((entry[0], entry[1])):
            hs = hs.to(device=dev, dtype=q.dtype).reshape(Fl * N, C)
            native.gather_rows(hs, s_idx, Fl * N, count=s_count, out=g[half * cap:half * cap + Fl * N])
        if plan is not None:
            kv = self._proj(g, plan[2], kv=True)
            return kv[:, :C], kv[:, C:], cap, ranges
        return attn.to_k(g), attn.to_v(g), cap, ranges

    def _attn_read(self, attn, entry, q, k, v, o, N, heads, cm, plan=None):
        """Read mode: keys = id_bank rows (all of them for early steps, :94-96; the sampled ones for the consistent
        branch, mask[F*N:], :106-108) + the current frame, per CFG half.  The bank is K/V source A, the current
        projection source B; nothing is concatenated."""
        Fl = self.id_length
        R = q.shape[0] // (2 * N)          # generated frames in this call (1 unless batched_read)
        native.flush_batch()               # entry.kv() may project bank rows with torch: nothing deferred before it
        own = dict(k_b=k, v_b=v, b_group_rows=R * N, cb=(0, N, N))
        if (not entry.has_kv() and cm is not None and self.kv_gather == "pre" and cm.shared_sample
                and entry[0] is not None and entry[1] is not None
                and tuple(entry[0].shape) == (Fl, N, q.shape[1]) and tuple(entry[1].shape) == (Fl, N, q.shape[1])
                and getattr(attn.to_k, "bias", None) is None and getattr(attn.to_v, "bias", None) is None):
            k_s, v_s, cap, ranges = self._project_sampled_bank(attn, entry, plan, cm, N, q)
            native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=R, n_q=N,
                            k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=ranges, range_base=Fl, range_step=0, **own)
            return
        kb, vb = entry.kv(attn, device=q.device)
        if kb.shape[0] != 2 * Fl * N or kb.shape[1] != q.shape[1]:
            raise ValueError(f"id_bank entry has K/V of shape {tuple(kb.shape)}, expected {(2 * Fl * N, q.shape[1])}")
        if kb.dtype != q.dtype:
            kb, vb = kb.to(q.dtype), vb.to(q.dtype)
        # every generated frame r attends the same bank keys + its own block [r*N, (r+1)*N) of the current projection
        if cm is None:
            native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=R, n_q=N,
                            k_a=kb, v_a=vb, a_group_rows=Fl * N, ca=(0, 0, Fl * N), **own)
        elif self.kv_gather == "pre" and cm.shared_sample:
            s_idx, s_count, ranges = cm.sample_list(q.device)
            k_s, v_s, cap = native.gather_kv(kb, vb, Fl * N, 2, s_idx, s_count, Fl * N)
            native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=R, n_q=N,
                            k_a=k_s, v_a=v_s, a_group_rows=cap, ranges=ranges, range_base=Fl, range_step=0, **own)
        else:
            idx, counts = cm.lists(q.device)
            native.attn_fwd(q, o, heads=heads, n_groups=2, n_frames=R, n_q=N,
                            k_a=kb, v_a=vb, a_group_rows=Fl * N,
                            idx=idx, counts=counts, list_base=Fl, list_step=0, g_adjust=-N, **own)


def set_bank_store(mode: str, cls=SpatialAttnProcessor2_0) -> None:
    if mode not in STORE_MODES:
        raise ValueError(f"bank_store must be one of {STORE_MODES}")
    cls.bank_store = mode
