"""ctypes binding of libgpmap_b200.so (declared in include/gpmap_b200.h).

The library is the product: there is no CPU fallback.  Loading fails loudly when the shared object
is missing, and every call fails loudly when no sm_100a device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpmap_b200.so")


class GpmGrid(C.Structure):
    _fields_ = [("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double), ("y1", C.c_double),
                ("t", C.c_double), ("gx", C.c_int32), ("gy", C.c_int32)]


COV_FULL, COV_LOWER = 0, 1
PREDICT_MEAN, PREDICT_VAR, PREDICT_ADD_NOISE = 1, 2, 4

_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t

# name -> (restype, argtypes); mirrors include/gpmap_b200.h one to one
SIGNATURES = {
    "gpm_version": (C.c_int, []),
    "gpm_last_error": (C.c_char_p, []),
    "gpm_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "gpm_destroy": (C.c_int, [_vp]),
    "gpm_sm_count": (C.c_int, [_vp]),
    "gpm_launch_count": (C.c_longlong, []),
    "gpm_cov": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _i32, _vp]),
    "gpm_cross_cov": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, C.POINTER(GpmGrid), _i64, _i64,
                                _vp, _i64, _vp]),
    "gpm_potrf_workspace_bytes": (_sz, [_i64]),
    "gpm_potrf": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "gpm_solve_lml": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gpm_predict_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "gpm_predict": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _vp, _vp, _i32,
                              _vp, C.POINTER(GpmGrid), _i64, _i64, _vp, _vp, _vp, _sz, _i32, _vp]),
    "gpm_fit_batched_workspace_bytes": (_sz, [_i64, _i64]),
    "gpm_fit_batched": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, C.POINTER(C.c_double), _i64,
                                  _vp, _vp, _vp, _vp, _vp]),
    "gpm_lml_grad_workspace_bytes": (_sz, [_i64]),
    "gpm_lml_grad": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gpm_kmeans_assign": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and set the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gaussianprocesspathmodelling_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library needs it)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GpmError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().gpm_last_error().decode("utf-8", "replace")
        raise GpmError(f"{what} failed with code {rc}: {msg}")


_handles = {}


def handle(device_index: int):
    """Per-device library handle (helper stream + events), created on first use."""
    h = _handles.get(device_index)
    if h is None:
        lib = load()
        hv = _vp()
        check(lib.gpm_create(C.byref(hv), int(device_index)), "gpm_create")
        h = hv
        _handles[device_index] = h
    return h


def theta_array(theta):
    arr = (C.c_double * len(theta))(*[float(v) for v in theta])
    return arr
