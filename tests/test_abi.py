"""CPU checks of the C-ABI boundary: libt2h.so loads and exports every symbol
include/t2h.h declares, the ctypes struct mirrors the C struct, and argument
validation fails loudly (no compute calls — there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from text2human_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "t2h.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(t2h_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"libt2h.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in text2human_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.t2h_version() == 201
    assert isinstance(lib.t2h_last_error(), bytes)


def test_struct_layout_matches_header():
    code = r'''
#include "t2h.h"
#include <stdio.h>
#include <stddef.h>
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(t2h_tapgemm_params),
  offsetof(t2h_tapgemm_params,b), offsetof(t2h_tapgemm_params,ntaps), offsetof(t2h_tapgemm_params,d),
  offsetof(t2h_tapgemm_params,alpha), offsetof(t2h_tapgemm_params,gn_cpg),
  sizeof(t2h_conv_wgrad_params), offsetof(t2h_conv_wgrad_params,x), offsetof(t2h_conv_wgrad_params,ntaps),
  offsetof(t2h_conv_wgrad_params,dw), offsetof(t2h_conv_wgrad_params,k_split)); return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    P = _lib.TapGemmParams
    W = _lib.ConvWgradParams
    want = [ctypes.sizeof(P), P.b.offset, P.ntaps.offset, P.d.offset, P.alpha.offset, P.gn_cpg.offset,
            ctypes.sizeof(W), W.x.offset, W.ntaps.offset, W.dw.offset, W.k_split.offset]
    assert got == want


def test_argument_validation_reports_errors():
    lib = _lib.load()
    p = _lib.TapGemmParams()  # all zero: null operands
    rc = lib.t2h_tapgemm(ctypes.byref(p), None)
    assert rc == -1
    assert b"null operand" in lib.t2h_last_error()
    with pytest.raises(_lib.T2HError):
        _lib.check(rc)
    assert lib.t2h_gn_stats(None, None, 1, 1, 32, 32, None) == -1
    assert lib.t2h_vq_search(None, None, None, 1, 1, 1, 4, 1, 1, 1, 1, None, None, None, None, None, None,
                             None, 0, None) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from text2human_b200 import ops
    with pytest.raises(_lib.T2HError):
        ops.nchw_to_nhwc(torch.zeros(1, 8, 4, 4))
