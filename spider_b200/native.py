"""ctypes binding of ``libcsa_b200.so`` (the C ABI declared in ``include/csa_b200.h``).

The library is the product's only compute path for the consistent-self-attention hot path: there is no CPU or
PyTorch fallback.  If the shared object is missing or the device is not an sm_100 part, calls raise
``CsaNativeError`` loudly instead of degrading.

PyTorch is used here only for device memory (``tensor.data_ptr()``) and for the current CUDA stream.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint8, c_uint32, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSA_B200_LIB: developer knob to load an experimental build of the same ABI (kernel tuning sweeps); default in-tree
LIB_PATH = os.environ.get("CSA_B200_LIB") or os.path.join(_HERE, "libcsa_b200.so")

CSA_ABI_VERSION = 9
CSA_DTYPE_F16 = 0
CSA_DTYPE_BF16 = 1
CSA_TILE = 128
CSA_HEAD_DIM = 64
CSA_MAX_PEERS = 8

# every symbol include/csa_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = (
    "csa_abi_version",
    "csa_last_error",
    "csa_device_supported",
    "csa_debug_stuck",
    "csa_debug_set_trace",
    "csa_compact_rows",
    "csa_validate_mask",
    "csa_gather_rows",
    "csa_sample_ranges",
    "csa_gather_kv",
    "csa_attn_fwd",
    "csa_attn_workspace_bytes",
    "csa_debug_last_launch",
    "csa_peer_scatter_kv",
    "csa_peer_signal",
    "csa_peer_signal_ex",
    "csa_epoch_advance",
    "csa_enable_peer_access",
    "csa_ipc_export",
    "csa_ipc_open",
    "csa_ipc_close",
    "csa_linear",
    "csa_gemm",
    "csa_gemm_supported",
    "csa_sample_positions",
    "csa_run_batch",
)


class CsaNativeError(RuntimeError):
    """Raised when libcsa_b200.so is missing, mismatched, or a call into it fails."""


class CsaAttnArgs(ctypes.Structure):
    """Mirror of ``csa_attn_args_t`` (include/csa_b200.h) — field order and types must match exactly."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("dtype", c_int32),
        ("head_dim", c_int32),
        ("heads", c_int32),
        ("n_groups", c_int32),
        ("n_frames", c_int32),
        ("n_q", c_int32),
        ("scale", c_float),
        ("q", c_void_p),
        ("o", c_void_p),
        ("q_ld", c_int64),
        ("o_ld", c_int64),
        ("k_a", c_void_p),
        ("v_a", c_void_p),
        ("a_ld", c_int64),
        ("a_rows", c_int64),
        ("a_group_rows", c_int32),
        ("_pad0", c_int32),
        ("k_b", c_void_p),
        ("v_b", c_void_p),
        ("b_ld", c_int64),
        ("b_rows", c_int64),
        ("b_group_rows", c_int32),
        ("_pad1", c_int32),
        ("idx", c_void_p),
        ("counts", c_void_p),
        ("idx_stride", c_int64),
        ("list_base", c_int32),
        ("list_step", c_int32),
        ("g_adjust", c_int32),
        ("ca_start", c_int32),
        ("ca_step", c_int32),
        ("ca_len", c_int32),
        ("cb_start", c_int32),
        ("cb_step", c_int32),
        ("cb_len", c_int32),
        ("max_ctas", c_int32),
        ("flags", c_int32),
        ("ranges", c_void_p),
        ("range_base", c_int32),
        ("range_step", c_int32),
        ("workspace", c_void_p),
        ("workspace_bytes", c_int64),
        ("ready", c_void_p),
        ("ready_epoch", c_uint32),
        ("ready_n", c_int32),
        ("ready_bounds", c_int32 * (CSA_MAX_PEERS + 1)),
        ("ready_frames_per_peer", c_int32),
        ("_pad2", c_int32),
        ("epoch_base", c_void_p),
        ("done_dst", c_void_p * CSA_MAX_PEERS),
        ("done_counter", c_void_p),
        ("peer_self", c_int32),
        ("_pad3", c_int32),
    ]


class CsaPeerScatterArgs(ctypes.Structure):
    """Mirror of ``csa_peer_scatter_args_t`` (include/csa_b200.h)."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("n_peers", c_int32),
        ("self_", c_int32),
        ("row_bytes", c_int32),
        ("k", c_void_p),
        ("v", c_void_p),
        ("ld_bytes", c_int64),
        ("idx", c_void_p),
        ("count", c_int32),
        ("dst_row0", c_int32),
        ("k_dst", c_void_p * CSA_MAX_PEERS),
        ("v_dst", c_void_p * CSA_MAX_PEERS),
        ("dst_ld_bytes", c_int64),
        ("ready", c_void_p * CSA_MAX_PEERS),
        ("epoch", c_uint32),
        ("done_epoch", c_uint32),
        ("done", c_void_p),
        ("counter", c_void_p),
        ("ranges", c_void_p),
        ("frames_per_peer", c_int32),
        ("idx_adjust", c_int32),
        ("epoch_base", c_void_p),
    ]


class CsaLinearArgs(ctypes.Structure):
    """Mirror of ``csa_linear_args_t``."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("dtype", c_int32),
        ("m", c_int64),
        ("n", c_int64),
        ("k", c_int64),
        ("x", c_void_p),
        ("ldx", c_int64),
        ("w", c_void_p),
        ("ldw", c_int64),
        ("bias", c_void_p),
        ("y", c_void_p),
        ("ldy", c_int64),
        ("workspace", c_void_p),
        ("workspace_bytes", c_int64),
    ]


class CsaPeerExchange(ctypes.Structure):
    """Mirror of ``csa_peer_exchange_t``."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("n_peers", c_int32),
        ("self_", c_int32),
        ("_pad0", c_int32),
        ("k_dst", c_void_p * CSA_MAX_PEERS),
        ("v_dst", c_void_p * CSA_MAX_PEERS),
        ("dst_ld", c_int64),
        ("ready", c_void_p * CSA_MAX_PEERS),
        ("epoch", c_uint32),
        ("done_epoch", c_uint32),
        ("done", c_void_p),
        ("counter", c_void_p),
        ("epoch_base", c_void_p),
    ]


class CsaGemmArgs(ctypes.Structure):
    """Mirror of ``csa_gemm_args_t``."""

    _fields_ = [
        ("struct_size", c_uint32),
        ("dtype", c_int32),
        ("m", c_int64),
        ("n", c_int64),
        ("k", c_int64),
        ("x", c_void_p),
        ("ldx", c_int64),
        ("w", c_void_p),
        ("ldw", c_int64),
        ("bias", c_void_p),
        ("y", c_void_p),
        ("ldy", c_int64),
        ("alpha", c_float),
        ("_pad0", c_int32),
        ("scatter_pos", c_void_p),
        ("scatter_k", c_void_p),
        ("scatter_v", c_void_p),
        ("scatter_ld", c_int64),
        ("scatter_group_rows", c_int32),
        ("scatter_dst_group_rows", c_int32),
        ("split_col", c_int32),
        ("scatter_col0", c_int32),
        ("y2", c_void_p),
        ("ldy2", c_int64),
        ("y_split", c_int32),
        ("_pad1", c_int32),
        ("exchange", c_void_p),
    ]


class CsaGatherKvArgs(ctypes.Structure):
    """Mirror of ``csa_gather_kv_args_t`` (the arguments of csa_gather_kv, for csa_run_batch)."""

    _fields_ = [
        ("k", c_void_p),
        ("v", c_void_p),
        ("ld_bytes", c_int64),
        ("group_rows", c_int32),
        ("n_groups", c_int32),
        ("s_idx", c_void_p),
        ("s_count", c_void_p),
        ("max_rows", c_int32),
        ("_pad0", c_int32),
        ("k_out", c_void_p),
        ("v_out", c_void_p),
        ("out_ld_bytes", c_int64),
        ("out_group_rows", c_int32),
        ("row_bytes", c_int32),
    ]


class CsaPeerSignalArgs(ctypes.Structure):
    """Mirror of ``csa_peer_signal_args_t``."""

    _fields_ = [
        ("done", c_void_p * CSA_MAX_PEERS),
        ("n_peers", c_int32),
        ("self_", c_int32),
        ("epoch", c_uint32),
        ("_pad0", c_uint32),
        ("epoch_base", c_void_p),
    ]


class CsaEpochAdvanceArgs(ctypes.Structure):
    """Mirror of ``csa_epoch_advance_args_t``."""

    _fields_ = [("epoch_base", c_void_p), ("delta", c_uint32), ("_pad0", c_uint32)]


class CsaCall(ctypes.Structure):
    """Mirror of ``csa_call_t``."""

    _fields_ = [("kind", c_int32), ("_pad0", c_int32), ("args", c_void_p)]


CSA_CALL_LINEAR, CSA_CALL_ATTN, CSA_CALL_GATHER_KV, CSA_CALL_PEER_SCATTER, CSA_CALL_PEER_SIGNAL, \
    CSA_CALL_EVENT_RECORD, CSA_CALL_EPOCH_ADVANCE, CSA_CALL_GEMM = 1, 2, 3, 4, 5, 6, 7, 8

CSA_ATTN_NO_SPLIT = 1
CSA_ATTN_B_FIRST = 2


_lib: Optional[ctypes.CDLL] = None

# Launch accounting (how many of OUR kernels were launched, by entry point) and optional CUDA-event timing of the
# attention launches on the launching stream; both are read by bench.py.
LAUNCHES = {"csa_attn_fwd": 0, "csa_compact_rows": 0, "csa_validate_mask": 0, "csa_gather_rows": 0,
            "csa_sample_ranges": 0, "csa_gather_kv": 0, "csa_peer_scatter_kv": 0, "csa_peer_signal": 0,
            "csa_linear": 0, "csa_epoch_advance": 0, "csa_gemm": 0, "csa_sample_positions": 0}
ATTN_EVENTS: Optional[list] = None   # when a list: (start_event, end_event, n_groups, n_frames, n_q, heads) appended


def reset_launch_counters() -> None:
    for k in LAUNCHES:
        LAUNCHES[k] = 0


def load() -> ctypes.CDLL:
    """Load the shared library once and declare the prototypes.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CsaNativeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C spider_b200/csrc`).  There is no CPU fallback for this path."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the box
        raise CsaNativeError(f"cannot load {LIB_PATH}: {e}") from e

    lib.csa_abi_version.restype = c_int32
    lib.csa_abi_version.argtypes = []
    lib.csa_last_error.restype = c_char_p
    lib.csa_last_error.argtypes = []
    lib.csa_device_supported.restype = c_int32
    lib.csa_device_supported.argtypes = [c_int32]
    lib.csa_debug_stuck.restype = c_int32
    lib.csa_debug_stuck.argtypes = [POINTER(c_uint32)]
    lib.csa_debug_set_trace.restype = c_int32
    lib.csa_debug_set_trace.argtypes = [c_void_p]
    lib.csa_compact_rows.restype = c_int32
    lib.csa_compact_rows.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64,
                                     c_void_p, c_void_p]
    lib.csa_validate_mask.restype = c_int32
    lib.csa_validate_mask.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.csa_gather_rows.restype = c_int32
    lib.csa_gather_rows.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                    c_int64, c_int32, c_void_p]
    lib.csa_sample_ranges.restype = c_int32
    lib.csa_sample_ranges.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    lib.csa_gather_kv.restype = c_int32
    lib.csa_gather_kv.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                  c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]
    lib.csa_attn_fwd.restype = c_int32
    lib.csa_attn_fwd.argtypes = [POINTER(CsaAttnArgs), c_void_p]
    lib.csa_debug_last_launch.restype = c_int32
    lib.csa_debug_last_launch.argtypes = [POINTER(c_int32)]
    lib.csa_attn_workspace_bytes.restype = c_int64
    lib.csa_attn_workspace_bytes.argtypes = [c_int32]
    lib.csa_peer_scatter_kv.restype = c_int32
    lib.csa_peer_scatter_kv.argtypes = [POINTER(CsaPeerScatterArgs), c_void_p]
    lib.csa_peer_signal.restype = c_int32
    lib.csa_peer_signal.argtypes = [POINTER(c_void_p), c_int32, c_int32, c_uint32, c_void_p]
    lib.csa_peer_signal_ex.restype = c_int32
    lib.csa_peer_signal_ex.argtypes = [POINTER(CsaPeerSignalArgs), c_void_p]
    lib.csa_epoch_advance.restype = c_int32
    lib.csa_epoch_advance.argtypes = [c_void_p, c_uint32, c_void_p]
    lib.csa_enable_peer_access.restype = c_int32
    lib.csa_enable_peer_access.argtypes = [c_int32]
    lib.csa_ipc_export.restype = c_int32
    lib.csa_ipc_export.argtypes = [c_void_p, c_void_p, POINTER(c_int64)]
    lib.csa_ipc_open.restype = c_int32
    lib.csa_ipc_open.argtypes = [c_void_p, POINTER(c_void_p)]
    lib.csa_ipc_close.restype = c_int32
    lib.csa_ipc_close.argtypes = [c_void_p]
    lib.csa_linear.restype = c_int32
    lib.csa_linear.argtypes = [POINTER(CsaLinearArgs), c_void_p]
    lib.csa_gemm.restype = c_int32
    lib.csa_gemm.argtypes = [POINTER(CsaGemmArgs), c_void_p]
    lib.csa_gemm_supported.restype = c_int32
    lib.csa_gemm_supported.argtypes = [c_int64, c_int64, c_int64]
    lib.csa_sample_positions.restype = c_int32
    lib.csa_sample_positions.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
    lib.csa_run_batch.restype = c_int32
    lib.csa_run_batch.argtypes = [POINTER(CsaCall), c_int32, c_void_p, POINTER(c_int32)]

    v = lib.csa_abi_version()
    if v != CSA_ABI_VERSION:
        raise CsaNativeError(f"libcsa_b200.so ABI version {v} != expected {CSA_ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().csa_last_error().decode("utf-8", "replace")
        raise CsaNativeError(f"{what} failed (rc={rc}): {msg}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(t: torch.Tensor) -> int:
    """cudaStream_t of torch's current stream on the tensor's device.  ``torch.cuda.current_stream()`` builds a Stream
    object (~10 us); the raw accessor is ~0.3 us, which matters once the GPU work of a layer shrinks (multi-GPU)."""
    idx = t.device.index
    if _raw_stream is not None:
        return _raw_stream(idx if idx is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CsaNativeError(
                "consistent self-attention needs CUDA tensors on a B200; got a CPU tensor and there is no CPU "
                "fallback in the product path")


_checked_devices: set = set()
_NULL_CTX = contextlib.nullcontext()


def _on_device_of(t: torch.Tensor):
    """Context that makes ``t``'s device the current CUDA device for a native call.  The library launches on the
    CURRENT device (``<<<>>>``, cuBLASLt, cudaGetDevice for the SM count) but takes its stream from the tensor's
    device: with a pipeline on cuda:1 while cuda:0 is current the two would disagree — an illegal address without
    peer access, silently the wrong GPU with it.  PyTorch ops guard this themselves; so do we."""
    idx = t.device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL_CTX
    return torch.cuda.device(idx)


def ensure_device(device: torch.device) -> None:
    idx = device.index
    if idx in _checked_devices:      # fast path: an explicit index that was already checked
        return
    if idx is None:
        idx = torch.cuda.current_device()
        if idx in _checked_devices:
            return
    _check(load().csa_device_supported(idx), "csa_device_supported")
    _checked_devices.add(idx)


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return CSA_DTYPE_BF16
    if dt == torch.float16:
        return CSA_DTYPE_F16
    raise CsaNativeError(f"consistent self-attention kernels support fp16/bf16 only, got {dt}")


# ---------------------------------------------------------------------------------------------- batched issue
# A processor call is a fixed sequence of launches (projections, K/V gather or peer exchange, attention, output
# projection).  Between begin_batch() and flush_batch() the wrappers below that are batchable only BUILD their
# argument blocks; flush_batch() hands the whole sequence to csa_run_batch — one transition into the library per
# processor call.  Nothing else may be enqueued on the stream in between (torch kernels would overtake the batch):
# code that has to do so calls flush_batch() first.
class _BatchState(threading.local):
    """Per-thread deferred-launch state: two threads driving processors (a multi-pipeline server) never interleave
    their batches.  ``calls`` holds ``(kind, args, keep)`` — ``keep`` references every tensor whose pointer the
    argument block carries, so that nothing a deferred launch reads or writes can be freed and re-used before the
    batch is flushed."""

    def __init__(self):
        self.calls: Optional[list] = None
        self.stream: int = 0
        self.like: Optional[torch.Tensor] = None   # a tensor of the batch's device (device guard at flush)


_TLS = _BatchState()
_BATCH_NAMES = {CSA_CALL_LINEAR: "csa_linear", CSA_CALL_ATTN: "csa_attn_fwd", CSA_CALL_GATHER_KV: "csa_gather_kv",
                CSA_CALL_PEER_SCATTER: "csa_peer_scatter_kv", CSA_CALL_PEER_SIGNAL: "csa_peer_signal",
                CSA_CALL_EVENT_RECORD: "cudaEventRecord", CSA_CALL_EPOCH_ADVANCE: "csa_epoch_advance",
                CSA_CALL_GEMM: "csa_gemm"}


def begin_batch(like: torch.Tensor) -> None:
    if _TLS.calls is not None:
        flush_batch()
    _TLS.calls = []
    _TLS.stream = _stream_ptr(like)
    _TLS.like = like


def flush_batch() -> None:
    """Issue what has been collected (no-op when no batch is open) and close the batch."""
    calls, _TLS.calls = _TLS.calls, None
    if not calls:
        return
    n = len(calls)
    arr = (CsaCall * n)()
    for i, (kind, a, _keep) in enumerate(calls):
        arr[i].kind = kind
        arr[i].args = a if isinstance(a, int) else ctypes.addressof(a)
    failed = c_int32(-1)
    with (_on_device_of(_TLS.like) if _TLS.like is not None else _NULL_CTX):
        rc = load().csa_run_batch(arr, n, _TLS.stream, ctypes.byref(failed))
    if rc != 0:
        what = _BATCH_NAMES.get(calls[failed.value][0], "?") if 0 <= failed.value < n else "csa_run_batch"
        _check(rc, f"{what} (entry {failed.value} of a batch of {n})")


def abort_batch() -> None:
    """Drop an open batch without issuing it (error paths)."""
    _TLS.calls = None


def _issue(kind: int, a, direct, what: str, stream: int, like: Optional[torch.Tensor] = None, keep=()) -> None:
    """Launch now (with ``like``'s device current), or defer into the open batch (same stream only); ``keep`` = the
    tensors the argument block points into."""
    if _TLS.calls is not None and stream == _TLS.stream:
        _TLS.calls.append((kind, a, keep))
    else:
        with (_on_device_of(like) if like is not None else _NULL_CTX):
            _check(direct(stream), what)


_EVENT_POOL: list = []


def _pooled_event() -> "torch.cuda.Event":
    """A timing event whose cudaEvent_t exists (torch creates it lazily at the first record)."""
    if _EVENT_POOL:
        return _EVENT_POOL.pop()
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def record_event(ev: "torch.cuda.Event", stream: int) -> None:
    """Record a timing event through the library (CSA_CALL_EVENT_RECORD): on a stream that is being captured into a
    CUDA graph this makes an event-record NODE whose time can be read after a replay, which torch's own
    ``Event.record()`` (a plainly captured event) does not allow."""
    if _TLS.calls is not None and stream == _TLS.stream:
        _TLS.calls.append((CSA_CALL_EVENT_RECORD, ev.cuda_event, (ev,)))
        return
    arr = (CsaCall * 1)()
    arr[0].kind = CSA_CALL_EVENT_RECORD
    arr[0].args = ev.cuda_event
    failed = c_int32(-1)
    _check(load().csa_run_batch(arr, 1, stream, ctypes.byref(failed)), "cudaEventRecord")


def prepare_event_pool(n: int) -> None:
    while len(_EVENT_POOL) < n:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        _EVENT_POOL.append(e)


_LINEAR_WS: dict = {}


def _linear_workspace(device: torch.device, stream: int) -> torch.Tensor:
    """cuBLASLt scratch, one per (device, stream): launches on one stream are ordered and may share it, two streams
    (two pipelines of one server) must not."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ws = _LINEAR_WS.get((idx, stream))
    if ws is None:
        flush_batch()
        ws = _LINEAR_WS[(idx, stream)] = torch.empty(32 << 20, dtype=torch.uint8, device=device)
    return ws


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[M, N] = x[M, K] @ w[N, K].T (+ bias)`` (see csa_linear): nn.Linear's layouts, 16-bit in/out, fp32
    accumulation.  2-D operands with unit column stride; batchable."""
    _require_cuda(x, w)
    ensure_device(x.device)
    if x.dim() != 2 or w.dim() != 2 or x.stride(1) != 1 or w.stride(1) != 1 or x.shape[1] != w.shape[1] or \
            w.dtype != x.dtype:
        raise CsaNativeError("linear expects 2-D x (M, K) and w (N, K) of one 16-bit dtype, unit column stride")
    m, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    elif out.shape != (m, n) or out.stride(1) != 1 or out.dtype != x.dtype:
        raise CsaNativeError("linear: out must be (M, N) of x's dtype with unit column stride")
    a = CsaLinearArgs()
    a.struct_size = ctypes.sizeof(CsaLinearArgs)
    a.dtype = dtype_code(x.dtype)
    a.m, a.n, a.k = m, n, k
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    a.w, a.ldw = w.data_ptr(), w.stride(0)
    if bias is not None:
        if bias.dtype != x.dtype or bias.numel() != n or not bias.is_contiguous():
            raise CsaNativeError("linear: bias must be a contiguous (N,) tensor of x's dtype")
        a.bias = bias.data_ptr()
    a.y, a.ldy = out.data_ptr(), out.stride(0)
    ws = _linear_workspace(x.device, _stream_ptr(x))
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    lib = load()
    _issue(CSA_CALL_LINEAR, a, lambda st: lib.csa_linear(ctypes.byref(a), st), "csa_linear", _stream_ptr(x), x,
           keep=(x, w, bias, out, ws))
    LAUNCHES["csa_linear"] += 1
    return out


def gemm_supported(m: int, n: int, k: int) -> bool:
    """Shapes the hand-written GEMM takes (N % 128 == 0, K % 64 == 0); anything else stays with ``linear``."""
    return bool(load().csa_gemm_supported(m, n, k))


def gemm(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         alpha: float = 1.0, scatter=None, out2: Optional[torch.Tensor] = None, exchange: Optional[dict] = None):
    """``out[M, N] = alpha * x[M, K] @ w[N, K].T (+ bias)`` with the hand-written sm_100a GEMM (see csa_gemm), same
    layouts as ``linear``; batchable.  ``scatter = (pos, k_s, v_s, group_rows, dst_group_rows, split_col[, col0])``
    fuses the gather of the sampled key rows into the epilogue (w = [w_k; w_v], or [w_q; w_k; w_v] with ``col0 = C``):
    a row ``r`` with ``pos[r % group_rows] >= 0`` is also stored into ``k_s`` / ``v_s`` at row
    ``(r // group_rows) * dst_group_rows + pos``.  With ``out2`` the product is split by column: ``out`` takes the
    first ``out.shape[1]`` columns, ``out2`` the rest (q and K|V of a stacked-weight GEMM in two buffers); returns
    ``(out, out2)`` then.  ``exchange`` (a dict with the arguments of ``peer_scatter_kv``: k_dst, v_dst, ready, self,
    epoch, done, done_epoch, counter, epoch_base) turns the fused gather into the multi-GPU exchange: the sampled rows
    go to every GPU's K[S] / V[S] buffer and the launch raises this GPU's arrival flag there (csa_peer_exchange_t);
    ``scatter = (pos, None, None, group_rows, 0, split_col[, col0])`` then."""
    _require_cuda(x, w)
    ensure_device(x.device)
    if x.dim() != 2 or w.dim() != 2 or x.stride(1) != 1 or w.stride(1) != 1 or x.shape[1] != w.shape[1] or \
            w.dtype != x.dtype:
        raise CsaNativeError("gemm expects 2-D x (M, K) and w (N, K) of one 16-bit dtype, unit column stride")
    m, k = x.shape
    n = w.shape[0]
    if out2 is not None:
        if out is None or out.dim() != 2 or out2.dim() != 2 or out.shape[0] != m or out2.shape[0] != m or \
                out.shape[1] + out2.shape[1] != n or out.stride(1) != 1 or out2.stride(1) != 1 or \
                out.dtype != x.dtype or out2.dtype != x.dtype:
            raise CsaNativeError("gemm: out and out2 must be (M, n1) and (M, N - n1) of x's dtype, unit column stride")
    elif out is None:
        out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    elif out.shape != (m, n) or out.stride(1) != 1 or out.dtype != x.dtype:
        raise CsaNativeError("gemm: out must be (M, N) of x's dtype with unit column stride")
    a = CsaGemmArgs()
    a.struct_size = ctypes.sizeof(CsaGemmArgs)
    a.dtype = dtype_code(x.dtype)
    a.m, a.n, a.k = m, n, k
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    a.w, a.ldw = w.data_ptr(), w.stride(0)
    if bias is not None:
        if bias.dtype != x.dtype or bias.numel() != n or not bias.is_contiguous():
            raise CsaNativeError("gemm: bias must be a contiguous (N,) tensor of x's dtype")
        a.bias = bias.data_ptr()
    a.y, a.ldy = out.data_ptr(), out.stride(0)
    if out2 is not None:
        a.y2, a.ldy2, a.y_split = out2.data_ptr(), out2.stride(0), out.shape[1]
    a.alpha = alpha
    xch = None
    if exchange is not None:
        if scatter is None:
            raise CsaNativeError("gemm: the exchange needs scatter = (pos, None, None, group_rows, 0, split_col[, col0])")
        pos, _, _, group_rows, _, split_col = scatter[:6]
        if pos.dtype != torch.int32 or not pos.is_contiguous() or pos.numel() < group_rows or m != group_rows:
            raise CsaNativeError("gemm: exchange needs int32 positions for exactly the M rows of x (one group)")
        k_dst, v_dst, ready = exchange["k_dst"], exchange["v_dst"], exchange["ready"]
        npeer = len(k_dst)
        if not (1 <= npeer <= CSA_MAX_PEERS) or len(v_dst) != npeer or len(ready) != npeer:
            raise CsaNativeError(f"gemm: exchange takes 1..{CSA_MAX_PEERS} peers, one K, V and flag buffer each")
        xch = CsaPeerExchange()
        xch.struct_size = ctypes.sizeof(CsaPeerExchange)
        xch.n_peers, xch.self_ = npeer, exchange["self"]
        ld = None
        for r in range(npeer):
            kd, vd = k_dst[r], v_dst[r]
            if kd.dim() != 2 or kd.stride(1) != 1 or kd.dtype != x.dtype or vd.shape != kd.shape or \
                    vd.stride(0) != kd.stride(0) or (ld is not None and kd.stride(0) != ld):
                raise CsaNativeError("gemm: exchange buffers must be 2-D, of x's dtype and one row stride")
            ld = kd.stride(0)
            xch.k_dst[r], xch.v_dst[r], xch.ready[r] = kd.data_ptr(), vd.data_ptr(), ready[r].data_ptr()
        xch.dst_ld = ld
        xch.epoch, xch.done_epoch = exchange["epoch"], exchange["done_epoch"] & 0xffffffff
        xch.done, xch.counter = exchange["done"].data_ptr(), exchange["counter"].data_ptr()
        if exchange.get("epoch_base") is not None:
            xch.epoch_base = exchange["epoch_base"].data_ptr()
        a.scatter_pos, a.scatter_group_rows, a.split_col = pos.data_ptr(), group_rows, split_col
        a.scatter_col0 = scatter[6] if len(scatter) > 6 else 0
        a.exchange = ctypes.addressof(xch)
    elif scatter is not None:
        pos, k_s, v_s, group_rows, dst_group_rows, split_col = scatter[:6]
        a.scatter_col0 = scatter[6] if len(scatter) > 6 else 0
        if pos.dtype != torch.int32 or not pos.is_contiguous() or pos.numel() < group_rows or \
                k_s.dim() != 2 or k_s.stride(1) != 1 or v_s.shape != k_s.shape or v_s.stride(0) != k_s.stride(0) or \
                k_s.dtype != x.dtype or k_s.shape[0] < (m // group_rows) * dst_group_rows:
            raise CsaNativeError("gemm: bad scatter buffers")
        a.scatter_pos, a.scatter_k, a.scatter_v = pos.data_ptr(), k_s.data_ptr(), v_s.data_ptr()
        a.scatter_ld, a.scatter_group_rows = k_s.stride(0), group_rows
        a.scatter_dst_group_rows, a.split_col = dst_group_rows, split_col
    lib = load()
    _issue(CSA_CALL_GEMM, a, lambda st: lib.csa_gemm(ctypes.byref(a), st), "csa_gemm", _stream_ptr(x), x,
           keep=(x, w, bias, out, out2, scatter, xch, exchange))
    LAUNCHES["csa_gemm"] += 1
    return out if out2 is None else (out, out2)


def sample_positions(s_idx: torch.Tensor, s_count: torch.Tensor, n_cols: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pos [n_cols] int32: index of every column in the ascending list ``s_idx`` (or -1); see csa_sample_positions."""
    flush_batch()
    _require_cuda(s_idx, s_count)
    ensure_device(s_idx.device)
    if out is None:
        out = torch.empty((n_cols,), dtype=torch.int32, device=s_idx.device)
    with _on_device_of(s_idx):
        rc = load().csa_sample_positions(s_idx.data_ptr(), s_count.data_ptr(), n_cols, out.data_ptr(),
                                         _stream_ptr(s_idx))
    _check(rc, "csa_sample_positions")
    LAUNCHES["csa_sample_positions"] += 1
    return out


def idx_stride_for(n_cols: int) -> int:
    return (n_cols + CSA_TILE - 1) // CSA_TILE * CSA_TILE


def compact_rows(mask_rows: torch.Tensor, n_rows: int, n_cols: int, row_stride: int, block_n: int = 0,
                 limit_cols: int = 0, idx: Optional[torch.Tensor] = None,
                 counts: Optional[torch.Tensor] = None):
    """Compact boolean rows into ascending int32 index lists (see csa_compact_rows).

    ``mask_rows`` is a bool/uint8 CUDA tensor whose first element is row 0, column 0; rows are ``row_stride``
    bytes apart (0 = the same vector for every row).  Returns ``(idx [n_rows, stride] int32, counts [n_rows])``.
    """
    flush_batch()   # launches immediately: whatever was deferred before it goes first
    _require_cuda(mask_rows)
    ensure_device(mask_rows.device)
    if mask_rows.dtype not in (torch.bool, torch.uint8):
        raise CsaNativeError(f"mask must be bool/uint8, got {mask_rows.dtype}")
    stride = idx_stride_for(n_cols)
    if idx is None:
        idx = torch.empty((n_rows, stride), dtype=torch.int32, device=mask_rows.device)
    if counts is None:
        counts = torch.empty((n_rows,), dtype=torch.int32, device=mask_rows.device)
    with _on_device_of(mask_rows):
        rc = load().csa_compact_rows(mask_rows.data_ptr(), row_stride, n_rows, n_cols, block_n, limit_cols,
                                     idx.data_ptr(), idx.stride(0), counts.data_ptr(), _stream_ptr(mask_rows))
    _check(rc, "csa_compact_rows")
    LAUNCHES["csa_compact_rows"] += 1
    return idx, counts


def validate_mask(mask: torch.Tensor, block_n: int) -> torch.Tensor:
    """Returns a device int32 scalar: number of 16-byte words that differ from their block's first row."""
    flush_batch()   # launches immediately: whatever was deferred before it goes first
    _require_cuda(mask)
    ensure_device(mask.device)
    if mask.dim() != 2 or mask.stride(1) != 1:
        raise CsaNativeError("validate_mask expects a 2-D mask with unit column stride")
    n_bad = torch.zeros((1,), dtype=torch.int32, device=mask.device)
    with _on_device_of(mask):
        rc = load().csa_validate_mask(mask.data_ptr(), mask.stride(0), mask.shape[0], mask.shape[1], block_n,
                                      n_bad.data_ptr(), _stream_ptr(mask))
    _check(rc, "csa_validate_mask")
    LAUNCHES["csa_validate_mask"] += 1
    return n_bad


def gather_rows(src: torch.Tensor, idx: torch.Tensor, max_rows: int, row_base: int = 0,
                count: Optional[torch.Tensor] = None, count_adjust: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = src[row_base + idx[i]] for i < min(count + adjust, max_rows).  src is 2-D, rows contiguous."""
    flush_batch()   # launches immediately: whatever was deferred before it goes first
    _require_cuda(src, idx)
    ensure_device(src.device)
    if src.dim() != 2 or src.stride(1) != 1:
        raise CsaNativeError("gather_rows expects a 2-D source with unit column stride")
    es = src.element_size()
    if out is None:
        out = torch.empty((max_rows, src.shape[1]), dtype=src.dtype, device=src.device)
    with _on_device_of(src):
        rc = load().csa_gather_rows(src.data_ptr(), src.stride(0) * es, row_base, idx.data_ptr(),
                                    count.data_ptr() if count is not None else None, count_adjust, max_rows,
                                    out.data_ptr(), out.stride(0) * es, src.shape[1] * es, _stream_ptr(src))
    _check(rc, "csa_gather_rows")
    LAUNCHES["csa_gather_rows"] += 1
    return out


def sample_ranges(s_idx: torch.Tensor, s_count: torch.Tensor, block_n: int, n_frames: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ranges [(n_frames+1), 4] int32 (see csa_sample_ranges): the runs of the sampled list each frame attends."""
    flush_batch()   # launches immediately: whatever was deferred before it goes first
    _require_cuda(s_idx, s_count)
    ensure_device(s_idx.device)
    if out is not None:
        if out.dtype != torch.int32 or not out.is_contiguous() or out.numel() != (n_frames + 1) * 4:
            raise CsaNativeError("sample_ranges: out must be a contiguous int32 tensor of (n_frames + 1) * 4 elements")
        ranges = out
    else:
        ranges = torch.empty((n_frames + 1, 4), dtype=torch.int32, device=s_idx.device)
    with _on_device_of(s_idx):
        rc = load().csa_sample_ranges(s_idx.data_ptr(), s_count.data_ptr(), block_n, n_frames, ranges.data_ptr(),
                                      _stream_ptr(s_idx))
    _check(rc, "csa_sample_ranges")
    LAUNCHES["csa_sample_ranges"] += 1
    return ranges


def gather_kv(k: torch.Tensor, v: torch.Tensor, group_rows: int, n_groups: int, s_idx: torch.Tensor,
              s_count: torch.Tensor, max_rows: int):
    """Sampled K/V rows made contiguous per group (see csa_gather_kv).  Returns (k_s, v_s, out_group_rows) with
    k_s / v_s of shape (n_groups * out_group_rows, C); rows beyond the count are zero for one tile, then undefined."""
    _require_cuda(k, v, s_idx, s_count)
    ensure_device(k.device)
    if k.dim() != 2 or k.stride(1) != 1 or v.shape != k.shape or v.stride(0) != k.stride(0):
        raise CsaNativeError("gather_kv expects 2-D k/v of equal shape and row stride, unit column stride")
    es = k.element_size()
    out_group_rows = max_rows + CSA_TILE
    k_s = torch.empty((n_groups * out_group_rows, k.shape[1]), dtype=k.dtype, device=k.device)
    v_s = torch.empty_like(k_s)
    a = CsaGatherKvArgs()
    a.k, a.v, a.ld_bytes = k.data_ptr(), v.data_ptr(), k.stride(0) * es
    a.group_rows, a.n_groups = group_rows, n_groups
    a.s_idx, a.s_count, a.max_rows = s_idx.data_ptr(), s_count.data_ptr(), max_rows
    a.k_out, a.v_out, a.out_ld_bytes = k_s.data_ptr(), v_s.data_ptr(), k_s.stride(0) * es
    a.out_group_rows, a.row_bytes = out_group_rows, k.shape[1] * es
    lib = load()
    _issue(CSA_CALL_GATHER_KV, a,
           lambda st: lib.csa_gather_kv(a.k, a.v, a.ld_bytes, a.group_rows, a.n_groups, a.s_idx, a.s_count, a.max_rows,
                                        a.k_out, a.v_out, a.out_ld_bytes, a.out_group_rows, a.row_bytes, st),
           "csa_gather_kv", _stream_ptr(k), k, keep=(k, v, s_idx, s_count, k_s, v_s))
    LAUNCHES["csa_gather_kv"] += 1
    return k_s, v_s, out_group_rows


# Tail-split scratch of csa_attn_fwd: one zero-initialised buffer per (device, stream) — launches on one stream are
# ordered, so they can share it; the kernel leaves the header zero again.
_WORKSPACES: dict = {}


def attn_workspace(device: torch.device, stream_ptr: Optional[int] = None) -> torch.Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if stream_ptr is None:
        stream_ptr = _raw_stream(idx) if _raw_stream is not None else torch.cuda.current_stream(device).cuda_stream
    key = (idx, stream_ptr)
    ws = _WORKSPACES.get(key)
    if ws is None:
        with torch.cuda.device(device):
            n = int(load().csa_attn_workspace_bytes(0))
        if n <= 0:
            raise CsaNativeError("csa_attn_workspace_bytes failed: " + (load().csa_last_error() or b"").decode())
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def attn_fwd(q: torch.Tensor, o: torch.Tensor, *, heads: int, n_groups: int, n_frames: int, n_q: int,
             k_a: Optional[torch.Tensor] = None, v_a: Optional[torch.Tensor] = None, a_group_rows: int = 0,
             k_b: Optional[torch.Tensor] = None, v_b: Optional[torch.Tensor] = None, b_group_rows: int = 0,
             idx: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None,
             list_base: int = -1, list_step: int = 0, g_adjust: int = 0,
             ca: tuple = (0, 0, 0), cb: tuple = (0, 0, 0), scale: Optional[float] = None,
             max_ctas: int = 0, ranges: Optional[torch.Tensor] = None, range_base: int = 0,
             range_step: int = 0, split=True, b_first: bool = False, ready: Optional[torch.Tensor] = None,
             ready_epoch: int = 0, ready_bounds=None, ready_peers: int = 0,
             ready_frames_per_peer: int = 0, epoch_base: Optional[torch.Tensor] = None,
             done=None, done_counter: Optional[torch.Tensor] = None, peer_self: int = 0) -> torch.Tensor:
    """Launch csa_attn_fwd on the current stream.  All matrices are 2-D ``(rows, heads*64)`` with unit column stride.
    ``done`` (list of the peers' release-flag tensors) + ``done_counter`` (local zero int32) + ``peer_self``: the
    launch itself publishes ``ready_epoch`` to the peers when it has finished reading the exchange buffers (what
    ``peer_signal`` does in a launch of its own)."""
    _require_cuda(q, o)
    ensure_device(q.device)
    for t in (q, o, k_a, v_a, k_b, v_b):
        if t is not None and (t.dim() != 2 or t.stride(1) != 1):
            raise CsaNativeError("attn_fwd expects 2-D (rows, heads*64) tensors with unit column stride")
    a = CsaAttnArgs()
    a.struct_size = ctypes.sizeof(CsaAttnArgs)
    a.dtype = dtype_code(q.dtype)
    a.head_dim = CSA_HEAD_DIM
    a.heads = heads
    a.n_groups = n_groups
    a.n_frames = n_frames
    a.n_q = n_q
    a.scale = float(scale) if scale is not None else CSA_HEAD_DIM ** -0.5
    a.q = q.data_ptr()
    a.o = o.data_ptr()
    a.q_ld = q.stride(0)
    a.o_ld = o.stride(0)
    if k_a is not None:
        if v_a is None or v_a.stride(0) != k_a.stride(0) or v_a.shape != k_a.shape or k_a.dtype != q.dtype:
            raise CsaNativeError("k_a and v_a must have the same shape, row stride and the dtype of q")
        a.k_a, a.v_a = k_a.data_ptr(), v_a.data_ptr()
        a.a_ld, a.a_rows, a.a_group_rows = k_a.stride(0), k_a.shape[0], a_group_rows
    if k_b is not None:
        if v_b is None or v_b.stride(0) != k_b.stride(0) or v_b.shape != k_b.shape or k_b.dtype != q.dtype:
            raise CsaNativeError("k_b and v_b must have the same shape, row stride and the dtype of q")
        a.k_b, a.v_b = k_b.data_ptr(), v_b.data_ptr()
        a.b_ld, a.b_rows, a.b_group_rows = k_b.stride(0), k_b.shape[0], b_group_rows
    if idx is not None:
        a.idx, a.counts, a.idx_stride = idx.data_ptr(), counts.data_ptr(), idx.stride(0)
    a.list_base, a.list_step, a.g_adjust = list_base, list_step, g_adjust
    a.ca_start, a.ca_step, a.ca_len = ca
    a.cb_start, a.cb_step, a.cb_len = cb
    a.max_ctas = max_ctas
    # split: True = let the library decide, False = whole units only, int k = force k pieces (tests)
    a.flags = CSA_ATTN_NO_SPLIT if not split else (0 if split is True else (int(split) & 0xff) << 8)
    if b_first:
        a.flags |= CSA_ATTN_B_FIRST
    if ready is not None:
        # rows [ready_bounds[r], ready_bounds[r+1]) of A are delivered by peer r (peer_scatter_kv on that GPU)
        # (or, with ready_frames_per_peer, the sampled rows of that many frames each: bounds taken from `ranges` on
        # the device — no host read-back)
        n = ready_peers if ready_frames_per_peer > 0 else len(ready_bounds) - 1
        if ready.dtype != torch.int32 or not ready.is_contiguous() or ready.numel() < n or not 1 <= n <= CSA_MAX_PEERS:
            raise CsaNativeError("ready must be a contiguous int32 tensor with one flag per peer (<= 8 peers)")
        a.ready, a.ready_epoch, a.ready_n = ready.data_ptr(), ready_epoch, n
        if epoch_base is not None:
            a.epoch_base = epoch_base.data_ptr()
        if ready_frames_per_peer > 0:
            a.ready_frames_per_peer = ready_frames_per_peer
        else:
            for i, b in enumerate(ready_bounds):
                a.ready_bounds[i] = b
        if done_counter is not None:
            if done is None or len(done) != n:
                raise CsaNativeError("attn_fwd: done must hold one release-flag tensor per peer")
            for r, t in enumerate(done):
                a.done_dst[r] = t.data_ptr()
            a.done_counter, a.peer_self = done_counter.data_ptr(), peer_self
    stream = _stream_ptr(q)
    if split:
        ws = attn_workspace(q.device, stream)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    if ranges is not None:
        if ranges.dtype != torch.int32 or not ranges.is_contiguous() or ranges.shape[-1] != 4:
            raise CsaNativeError("ranges must be a contiguous int32 tensor of shape (lists, 4)")
        a.ranges, a.range_base, a.range_step = ranges.data_ptr(), range_base, range_step
    lib = load()
    keep = (q, o, k_a, v_a, k_b, v_b, idx, counts, ranges, ready, epoch_base, done, done_counter)
    direct = lambda st: lib.csa_attn_fwd(ctypes.byref(a), st)   # noqa: E731
    if ATTN_EVENTS is not None:
        # kernel time of the attention launches on the launching stream (bench.py's roofline line)
        e0, e1 = _pooled_event(), _pooled_event()
        record_event(e0, stream)
        _issue(CSA_CALL_ATTN, a, direct, "csa_attn_fwd", stream, q, keep=keep)
        record_event(e1, stream)
        ATTN_EVENTS.append((e0, e1, n_groups, n_frames, n_q, heads))
    else:
        _issue(CSA_CALL_ATTN, a, direct, "csa_attn_fwd", stream, q, keep=keep)
    LAUNCHES["csa_attn_fwd"] += 1
    return o


def peer_scatter_kv(k: torch.Tensor, v: torch.Tensor, idx: torch.Tensor, count: int, dst_row0: int, k_dst, v_dst,
                    ready, self_index: int, epoch: int, done: torch.Tensor, done_epoch: int,
                    counter: torch.Tensor, ranges: Optional[torch.Tensor] = None, frames_per_peer: int = 0,
                    idx_adjust: int = 0, epoch_base: Optional[torch.Tensor] = None) -> None:
    """Store this GPU's sampled K/V rows into the S-ordered buffers of every GPU of the group and raise its arrival
    flag there (see csa_peer_scatter_kv).  ``k_dst / v_dst / ready`` are lists of tensors, one per GPU (peer memory;
    entry ``self_index`` is local); ``done`` and ``counter`` are local.  With ``ranges`` the geometry is read on the
    device (``idx`` = the whole sampled list, ``count`` = an upper bound, ``dst_row0`` ignored)."""
    _require_cuda(k, v, done, counter)
    ensure_device(k.device)
    n = len(k_dst)
    if not (1 <= n <= CSA_MAX_PEERS) or len(v_dst) != n or len(ready) != n:
        raise CsaNativeError(f"peer_scatter_kv: 1..{CSA_MAX_PEERS} peers, one K, V and flag buffer each")
    if k.dim() != 2 or k.stride(1) != 1 or v.shape != k.shape or v.stride(0) != k.stride(0):
        raise CsaNativeError("peer_scatter_kv expects 2-D k/v of equal shape and row stride, unit column stride")
    es = k.element_size()
    a = CsaPeerScatterArgs()
    a.struct_size = ctypes.sizeof(CsaPeerScatterArgs)
    a.n_peers, a.self_, a.row_bytes = n, self_index, k.shape[1] * es
    a.k, a.v, a.ld_bytes = k.data_ptr(), v.data_ptr(), k.stride(0) * es
    a.idx, a.count, a.dst_row0 = (idx.data_ptr() if count > 0 else None), count, dst_row0
    ld = None
    for r in range(n):
        kd, vd = k_dst[r], v_dst[r]
        if kd.dim() != 2 or kd.stride(1) != 1 or kd.shape[1] != k.shape[1] or kd.dtype != k.dtype or \
                vd.shape != kd.shape or vd.stride(0) != kd.stride(0) or (ld is not None and kd.stride(0) != ld):
            raise CsaNativeError("peer_scatter_kv: destination buffers must be 2-D, of k's width/dtype and one stride")
        if (0 if ranges is not None else dst_row0) + count > kd.shape[0]:
            raise CsaNativeError(f"peer_scatter_kv: rows [{dst_row0}, {dst_row0 + count}) exceed the buffer "
                                 f"({kd.shape[0]} rows)")
        ld = kd.stride(0)
        a.k_dst[r], a.v_dst[r], a.ready[r] = kd.data_ptr(), vd.data_ptr(), ready[r].data_ptr()
    a.dst_ld_bytes = ld * es
    # with epoch_base both are offsets to the device-resident base; done_epoch may then be <= 0 (two's complement)
    a.epoch, a.done_epoch = epoch, done_epoch & 0xffffffff
    a.done, a.counter = done.data_ptr(), counter.data_ptr()
    if epoch_base is not None:
        a.epoch_base = epoch_base.data_ptr()
    if ranges is not None:
        a.ranges, a.frames_per_peer, a.idx_adjust = ranges.data_ptr(), frames_per_peer, idx_adjust
    lib = load()
    _issue(CSA_CALL_PEER_SCATTER, a, lambda st: lib.csa_peer_scatter_kv(ctypes.byref(a), st), "csa_peer_scatter_kv",
           _stream_ptr(k), k, keep=(k, v, idx, k_dst, v_dst, ready, done, counter, ranges, epoch_base))
    LAUNCHES["csa_peer_scatter_kv"] += 1


def peer_signal(done, self_index: int, epoch: int, like: torch.Tensor,
                epoch_base: Optional[torch.Tensor] = None) -> None:
    """done[r][self_index] = epoch on every other GPU r, ordered after everything enqueued so far on the current
    stream of ``like``'s device (see csa_peer_signal)."""
    n = len(done)
    a = CsaPeerSignalArgs()
    for r, t in enumerate(done):
        a.done[r] = t.data_ptr()
    a.n_peers, a.self_, a.epoch = n, self_index, epoch
    if epoch_base is not None:
        a.epoch_base = epoch_base.data_ptr()
    lib = load()
    _issue(CSA_CALL_PEER_SIGNAL, a, lambda st: lib.csa_peer_signal_ex(ctypes.byref(a), st), "csa_peer_signal",
           _stream_ptr(like), like)
    LAUNCHES["csa_peer_signal"] += 1


def epoch_advance(epoch_base: torch.Tensor, delta: int) -> None:
    """``*epoch_base += delta`` on the current stream (see csa_epoch_advance): ends a step whose exchange calls
    used epochs relative to the device-resident base."""
    _require_cuda(epoch_base)
    a = CsaEpochAdvanceArgs()
    a.epoch_base, a.delta = epoch_base.data_ptr(), delta
    lib = load()
    _issue(CSA_CALL_EPOCH_ADVANCE, a, lambda st: lib.csa_epoch_advance(a.epoch_base, a.delta, st),
           "csa_epoch_advance", _stream_ptr(epoch_base), epoch_base)
    LAUNCHES["csa_epoch_advance"] += 1


def enable_peer_access(peer_device: int) -> None:
    _check(load().csa_enable_peer_access(peer_device), "csa_enable_peer_access")


def ipc_export(t: torch.Tensor):
    """(handle bytes, offset, nbytes) that lets another process of this box map ``t``'s memory (csa_ipc_export)."""
    _require_cuda(t)
    h = ctypes.create_string_buffer(64)
    off = c_int64(0)
    with torch.cuda.device(t.device):
        _check(load().csa_ipc_export(t.data_ptr(), h, ctypes.byref(off)), "csa_ipc_export")
    return bytes(h.raw), int(off.value), t.numel() * t.element_size()


class _RawDeviceMemory:
    """``__cuda_array_interface__`` view of mapped peer memory, so that torch can alias it as a uint8 tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


_IPC_OPEN: dict = {}   # (device index, handle bytes) -> mapped base: an allocation can be opened once per process


def ipc_import(handle, device: torch.device) -> torch.Tensor:
    """Map another process's allocation into ``device``'s address space; returns a uint8 tensor aliasing the exported
    bytes.  The mapping stays open for the life of the process (the exchange buffers live as long as the job)."""
    hbytes, off, nbytes = handle
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(idx):
        base = _IPC_OPEN.get((idx, hbytes))
        if base is None:
            out = c_void_p(0)
            _check(load().csa_ipc_open(hbytes, ctypes.byref(out)), "csa_ipc_open")
            base = _IPC_OPEN[(idx, hbytes)] = out.value
        return torch.as_tensor(_RawDeviceMemory(base + off, nbytes), device=torch.device("cuda", idx))


def last_launch() -> dict:
    """Work decomposition of this thread's last attention launch (csa_debug_last_launch)."""
    out = (c_int32 * 4)()
    _check(load().csa_debug_last_launch(out), "csa_debug_last_launch")
    return {"ctas": out[0], "whole_units": out[1], "split": out[2], "scheduled": out[3]}


def debug_stuck():
    out = (c_uint32 * 4)()
    if load().csa_debug_stuck(out):
        return tuple(out)
    return None
