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


def test_struct_layout_matches_c(tmp_path):
    """the header is plain C: gcc lays the argument structs out, and the ctypes mirrors must agree on size and on every field offset
    (a field added on one side only would shift everything behind it without any error at the call)"""
    import ctypes
    import shutil
    import subprocess
    assert ctypes.sizeof(L.ConvGeom) == 18 * 4
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    structs = {"avec_conv_geom": L.ConvGeom, "avec_gemm_args": L.GemmArgs, "avec_copy_job": L.CopyJob}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "avec_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'    printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'    printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["    return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, field, value = ln.split()
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "sizeof" else getattr(cls, field).offset
        assert int(value) == want, f"{cname}.{field}: C {value}, ctypes {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
    assert L.COPY_CHUNK == 4096        # AVEC_COPY_CHUNK of the header (chunk table granularity of avec_convert_multi)


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear_fwd(torch.zeros(4, 8), torch.zeros(4, 8))


def test_fused_dropout_policy_declines_before_any_launch():
    """ops.FUSE_DROPOUT: 0 never fuses, 2 (default) leaves GEMMs whose output is wider than twice the contraction to the separate
    dropout kernel - both decided on the host, before the library is called"""
    a = L.GemmArgs()
    prev = ops.FUSE_DROPOUT
    try:
        ops.FUSE_DROPOUT = 2
        a.N, a.K = 720, 180            # FFN first Linear: epilogue-bound
        assert ops._gemm_drop(a, (None, 0.1, 3)) is False
        ops.FUSE_DROPOUT = 0
        a.N, a.K = 180, 720
        assert ops._gemm_drop(a, (None, 0.1, 3)) is False
        assert a.drop_p == 0.0 and not a.drop_rng
    finally:
        ops.FUSE_DROPOUT = prev
    assert prev == int(os.environ.get("AVEC_FUSE_DROPOUT", "2"))
