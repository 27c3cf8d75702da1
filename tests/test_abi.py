"""CPU tests of the boundary: the shared library loads without a GPU and exports every symbol include/avec_b200.h
declares; the ctypes prototypes cover exactly that list; ops refuse CPU tensors (no silent fallback)."""
import os
import re

import pytest
import torch

from avec_b200 import _lib as L, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "avec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(avec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libavec_b200.so does not export {n}"
    assert sorted(L.PROTOTYPES) == names, "ctypes prototypes and header disagree"
    assert lib.avec_version() == 100
    assert lib.avec_strerror(-3) == b"combination not implemented"


def test_struct_layout_matches_c():
    # sizeof(avec_gemm_args) as laid out by the C compiler: 5 ints, ptr, 2 ll, ptr, 2 ll, int, geom(18 ints), ...
    import ctypes
    assert ctypes.sizeof(L.ConvGeom) == 18 * 4
    assert ctypes.sizeof(L.GemmArgs) % 8 == 0


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear_fwd(torch.zeros(4, 8), torch.zeros(4, 8))
