"""The C-ABI library loads on a CPU box and exports every symbol include/rsdf_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rsdf_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rsdf_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from rise_sdf_b200 import _lib
    from rise_sdf_b200.build import build
    build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rsdf_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == syms, "ctypes table and header disagree"
    L = _lib.lib()
    assert b"sm_100a" in L.rsdf_version()
    assert L.rsdf_error_string(-1) == b"rsdf: bad argument"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rise_sdf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith(".md"), f


def test_cuda_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from rise_sdf_b200 import nerfacc
    with pytest.raises(NotImplementedError):
        nerfacc.ray_marching(torch.zeros(4, 3), torch.ones(4, 3))
    with pytest.raises(NotImplementedError):
        nerfacc.render_weight_from_alpha(torch.rand(4), ray_indices=torch.zeros(4, dtype=torch.long), n_rays=1)
