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
    "gpm_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "gpm_get_option": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int)]),
    "gpm_cov": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _i32, _vp]),
    "gpm_cross_cov": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, C.POINTER(GpmGrid), _i64, _i64,
                                _vp, _i64, _vp]),
    "gpm_potrf_workspace_bytes": (_sz, [_i64]),
    "gpm_potrf": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "gpm_solve_lml": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gpm_fit": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "gpm_predict_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "gpm_predict": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _vp, _vp, _i32,
                              _vp, C.POINTER(GpmGrid), _i64, _i64, _vp, _vp, _vp, _sz, _i32, _vp]),
    "gpm_fit_batched_workspace_bytes": (_sz, [_vp, _i64, _i64]),
    "gpm_fit_batched": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, C.POINTER(C.c_double), _i64,
                                  _vp, _vp, _vp, _vp, _vp]),
    "gpm_lml_grad_workspace_bytes": (_sz, [_i64]),
    "gpm_lml_grad": (C.c_int, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double), _vp, _i64, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gpm_kmeans_assign": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gpm_kmeans_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "gpm_kmeans_lloyd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, C.c_double, _i32, _i32,
                                   _vp, _vp]),
    "gpm_kmeans_state": (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_double), _vp]),
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


_handles = {}          # (device index, stream id) -> handle
_options = {}          # option name -> value, applied to every handle (existing and future)
_MAX_HANDLES = 64


def handle(device_index: int, stream: int | None = None):
    """Library handle for (device, CUDA stream), created on first use.

    A handle is single-stream state (helper stream, events, the flags of the chained solves), so work issued on
    different streams -- possibly from different Python threads -- gets different handles and cannot race.
    ``stream`` defaults to torch's current stream on that device."""
    device_index = int(device_index)
    if stream is None:
        import torch
        stream = torch.cuda.current_stream(device_index).cuda_stream
    key = (device_index, int(stream))
    h = _handles.get(key)
    if h is None:
        lib = load()
        if len(_handles) >= _MAX_HANDLES:            # streams come and go: drop the oldest non-default-stream handle
            for old in list(_handles):
                if old[1] != 0:
                    lib.gpm_destroy(_handles.pop(old))
                    break
        hv = _vp()
        check(lib.gpm_create(C.byref(hv), device_index), "gpm_create")
        for name, value in _options.items():
            check(lib.gpm_set_option(hv, name.encode(), int(value)), f"gpm_set_option({name})")
        h = hv
        _handles[key] = h
    return h


def set_option(name: str, value: int):
    """Set a debug / comparison switch (see gpm_set_option in include/gpmap_b200.h) on every handle."""
    lib = load()
    _options[name] = int(value)
    for h in _handles.values():
        check(lib.gpm_set_option(h, name.encode(), int(value)), f"gpm_set_option({name})")


def get_option(name: str, device_index: int = 0) -> int:
    v = C.c_int(0)
    check(load().gpm_get_option(handle(device_index), name.encode(), C.byref(v)), f"gpm_get_option({name})")
    return v.value


class option:
    """Context manager: ``with _native.option("no_lookahead", 1): ...`` (restores the previous value)."""

    def __init__(self, name, value):
        self.name, self.value = name, int(value)

    def __enter__(self):
        self.prev = _options.get(self.name)
        self.prev_val = get_option(self.name) if self.prev is None else self.prev
        set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.prev_val)
        if self.prev is None:
            _options.pop(self.name, None)
        return False


def theta_array(theta):
    arr = (C.c_double * len(theta))(*[float(v) for v in theta])
    return arr
