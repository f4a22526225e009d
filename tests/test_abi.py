"""CPU: the C-ABI library loads, exports every symbol include/csa_b200.h declares, agrees with the ctypes mirror,
and rejects bad arguments before touching the GPU (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from spider_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "csa_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csa_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = native.load()
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/csa_b200.h but not exported by libcsa_b200.so"
    assert sorted(native.EXPORTED_SYMBOLS) == declared


STRUCTS = [("csa_attn_args_t", "CsaAttnArgs"), ("csa_peer_scatter_args_t", "CsaPeerScatterArgs"),
           ("csa_linear_args_t", "CsaLinearArgs"), ("csa_gather_kv_args_t", "CsaGatherKvArgs"),
           ("csa_peer_signal_args_t", "CsaPeerSignalArgs"), ("csa_call_t", "CsaCall")]


@pytest.mark.parametrize("c_name,py_name", STRUCTS)
def test_abi_version_and_struct_layout(tmp_path, c_name, py_name):
    lib = native.load()
    assert lib.csa_abi_version() == native.CSA_ABI_VERSION
    # sizeof/offsetof as the C compiler sees them == the ctypes mirror
    mirror = getattr(native, py_name)
    fields = [f[0] for f in mirror._fields_]
    c_field = lambda f: "self" if f == "self_" else f      # `self` is spelled self_ on the Python side
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            f'printf("%zu\\n", sizeof({c_name}));']
    for f in fields:
        prog.append(f'printf("%zu\\n", offsetof({c_name}, {c_field(f)}));')
    prog.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", str(c), "-o", str(exe)], check=True)
    vals = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(mirror)
    for f, off in zip(fields, vals[1:]):
        assert getattr(mirror, f).offset == off, f


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = native.load()
    assert lib.csa_attn_fwd(None, None) == -1
    assert b"null" in lib.csa_last_error()
    a = native.CsaAttnArgs()
    a.struct_size = 8
    assert lib.csa_attn_fwd(ctypes.byref(a), None) == -1
    assert b"ABI" in lib.csa_last_error()
    a.struct_size = ctypes.sizeof(native.CsaAttnArgs)
    a.head_dim = 128
    assert lib.csa_attn_fwd(ctypes.byref(a), None) == -2
    assert lib.csa_compact_rows(None, 0, 1, 1, 0, 0, None, 128, None, None) == -1
    assert lib.csa_validate_mask(None, 0, 1, 1, 1, None, None) == -1
    assert lib.csa_gather_rows(None, 0, 0, None, None, 0, 1, None, 0, 16, None) == -1
    # multi-GPU exchange, projections, batch runner: argument checks come before any CUDA call
    assert lib.csa_peer_scatter_kv(None, None) == -1
    ps = native.CsaPeerScatterArgs()
    ps.struct_size = ctypes.sizeof(native.CsaPeerScatterArgs)
    ps.n_peers, ps.self_ = 9, 0
    assert lib.csa_peer_scatter_kv(ctypes.byref(ps), None) == -1 and b"n_peers" in lib.csa_last_error()
    ps.n_peers, ps.self_, ps.row_bytes = 2, 2, 256
    assert lib.csa_peer_scatter_kv(ctypes.byref(ps), None) == -1
    assert lib.csa_peer_signal(None, 2, 0, 1, None) == -1
    assert lib.csa_ipc_export(None, None, None) == -1 and lib.csa_ipc_open(None, None) == -1
    assert lib.csa_linear(None, None) == -1
    la = native.CsaLinearArgs()
    la.struct_size = 4
    assert lib.csa_linear(ctypes.byref(la), None) == -1 and b"ABI" in lib.csa_last_error()
    la.struct_size = ctypes.sizeof(native.CsaLinearArgs)
    la.dtype, la.m, la.n, la.k, la.ldx, la.ldw, la.ldy = 1, 8, 8, 16, 8, 16, 8      # ldx < k
    assert lib.csa_linear(ctypes.byref(la), None) == -1 and b"bad sizes" in lib.csa_last_error()
    assert lib.csa_run_batch(None, 1, None, None) == -1
    calls = (native.CsaCall * 2)()
    calls[0].kind, calls[1].kind = native.CSA_CALL_LINEAR, 99
    failed = ctypes.c_int32(-5)
    assert lib.csa_run_batch(calls, 2, None, ctypes.byref(failed)) == -1 and failed.value == 0    # null linear args
    calls[0].kind = 99
    assert lib.csa_run_batch(calls, 2, None, ctypes.byref(failed)) == -1 and b"unknown call kind" in lib.csa_last_error()
    assert lib.csa_run_batch(calls, 0, None, ctypes.byref(failed)) == 0 and failed.value == -1


def test_product_path_has_no_cpu_fallback():
    """CPU tensors must raise, not silently compute somewhere else."""
    import torch

    from spider_b200 import SpatialAttnProcessor2_0
    from oracle.fake_diffusers import FakeAttention

    proc = SpatialAttnProcessor2_0(id_length=4)
    attn = FakeAttention(128, 2)
    with pytest.raises(native.CsaNativeError):
        proc(attn, torch.randn(8, 16, 128))
    with pytest.raises(native.CsaNativeError):
        native.compact_rows(torch.zeros(16, dtype=torch.bool), 1, 16, 0)


def test_product_does_not_import_oracle():
    """Nothing under spider_b200/ may import, call or link oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "spider_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libcsa_oracle" not in text, f
