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


def test_abi_version_and_struct_layout(tmp_path):
    lib = native.load()
    assert lib.csa_abi_version() == native.CSA_ABI_VERSION
    # sizeof/offsetof as the C compiler sees them == the ctypes mirror
    fields = [f[0] for f in native.CsaAttnArgs._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
            'printf("%zu\\n", sizeof(csa_attn_args_t));']
    for f in fields:
        prog.append(f'printf("%zu\\n", offsetof(csa_attn_args_t, {f}));')
    prog.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", str(c), "-o", str(exe)], check=True)
    vals = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(native.CsaAttnArgs)
    for f, off in zip(fields, vals[1:]):
        assert getattr(native.CsaAttnArgs, f).offset == off, f


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
