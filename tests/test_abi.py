"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes prototypes cover the header, and the product refuses to run without CUDA."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gpmap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gaussianprocesspathmodelling_b200 import build, _native
    build.build()
    lib = _native.load()
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gpmap_b200.h but not exported"
    assert lib.gpm_version() == 100


def test_ctypes_prototypes_cover_the_header():
    from gaussianprocesspathmodelling_b200 import _native
    assert sorted(_native.SIGNATURES) == header_functions()


def test_workspace_queries_need_no_gpu():
    from gaussianprocesspathmodelling_b200 import _native
    lib = _native.load()
    assert lib.gpm_potrf_workspace_bytes(4096) == 32 * 128 * 128 * 8
    assert lib.gpm_potrf_workspace_bytes(200) == 2 * 128 * 128 * 8
    assert lib.gpm_fit_batched_workspace_bytes(None, 3, 512) == 0      # depends on the handle's options: needs a handle
    # second centroid buffer + shifts + state, the rank / count tables and the ordered row buffer (P + 2k + 2 rows of 3n)
    assert lib.gpm_kmeans_workspace_bytes(100, 33, 3) >= (3 * 3 * 33 + 3 + 2) * 8 + (100 + 2 * 3 + 2) * 3 * 33 * 8
    assert lib.gpm_kmeans_workspace_bytes(100, 33, 3) % 16 == 0
    assert lib.gpm_potrf_workspace_bytes(0) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gaussianprocesspathmodelling_b200 as g
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.fit_gp([[0.0, 0.0], [1.0, 1.0]], [0.0, 1.0], lengthscale=1.0)
    from gaussianprocesspathmodelling_b200 import _native
    import ctypes as C
    h = C.c_void_p()
    assert _native.load().gpm_create(C.byref(h), 0) != 0          # no device -> error code, not a fallback


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaussianprocesspathmodelling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", ""), f"{f} mentions oracle/"
