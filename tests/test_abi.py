"""The C-ABI library loads on a CPU-only box and exports exactly what include/chefsi_b200.h
declares; with no GPU it refuses to create a context (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest

from sparc_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "chefsi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chefsi_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.chefsi_version()


def test_struct_layout_matches_header():
    from sparc_b200.problem import CHEFSI_MAX_FDN, ChefsiGridC, ChefsiNlocC
    assert C.sizeof(ChefsiGridC) == 8 * 4 + 8 * 4 + 15 * 8 * (CHEFSI_MAX_FDN + 1)
    assert C.sizeof(ChefsiNlocC) == 11 * 8
    assert C.sizeof(capi.ChefsiStats) == 8 + 3 * 8 + 10 * 4


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sparc_b200.chefsi import ChefsiContext
    with pytest.raises(capi.ChefsiError, match="no CPU fallback"):
        ChefsiContext(0)


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    import ast
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sparc_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for node in ast.walk(ast.parse(open(path).read())):
                    mods = []
                    if isinstance(node, ast.Import):
                        mods = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom):
                        mods = [node.module or ""]
                    assert not any(m.split(".")[0] == "oracle" for m in mods), path
            elif f.endswith((".cu", ".cuh", ".h", ".c", ".cpp")):
                for line in open(path):
                    if line.lstrip().startswith("#include"):
                        assert "oracle" not in line, (path, line)
    # and the shared library does not link against the checker
    import subprocess
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "sparc_ref" not in out
