"""ctypes access to oracle/_build/libcsa_oracle.so (oracle/compact_ref.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcsa_oracle.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = ctypes.CDLL(_LIB)
        lib.csa_ref_frame_row.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]
        lib.csa_ref_frame_row.restype = None
        lib.csa_ref_nonzero.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.csa_ref_nonzero.restype = ctypes.c_int
        lib.csa_ref_blocks_uniform.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int]
        lib.csa_ref_blocks_uniform.restype = ctypes.c_int
        lib.csa_ref_attention_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64]
        lib.csa_ref_attention_f32.restype = None
        _lib = lib
    return _lib


def frame_row(sample: np.ndarray, total_length: int, id_length: int, i: int) -> np.ndarray:
    sample = np.ascontiguousarray(sample.astype(np.uint8))
    n = sample.size // total_length
    out = np.empty(sample.size, dtype=np.uint8)
    load().csa_ref_frame_row(sample.ctypes.data, total_length, id_length, n, i, out.ctypes.data)
    return out


def nonzero(row: np.ndarray) -> np.ndarray:
    row = np.ascontiguousarray(row.astype(np.uint8))
    idx = np.empty(row.size, dtype=np.int32)
    c = load().csa_ref_nonzero(row.ctypes.data, row.size, idx.ctypes.data)
    return idx[:c].copy()


def blocks_uniform(mask: np.ndarray, block_n: int) -> bool:
    mask = np.ascontiguousarray(mask.astype(np.uint8))
    return bool(load().csa_ref_blocks_uniform(mask.ctypes.data, mask.strides[0], mask.shape[0], mask.shape[1],
                                               block_n))


def attention_f32(q: np.ndarray, k: np.ndarray, v: np.ndarray, idx: np.ndarray, scale: float) -> np.ndarray:
    """one head: q (nq, d), k/v (nk, d) float32, idx int32 key list -> (nq, d)"""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.empty_like(q)
    d = q.shape[1]
    load().csa_ref_attention_f32(q.ctypes.data, d, k.ctypes.data, v.ctypes.data, d, idx.ctypes.data, idx.size,
                                 q.shape[0], d, scale, out.ctypes.data, d)
    return out
