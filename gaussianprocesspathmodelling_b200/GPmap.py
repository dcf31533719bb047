"""Drop-in ``GPmap`` module: the reference's data model plus the B200-native GP hot path.

What is kept from ``/root/reference/GPmap.py`` (names, attributes, argument order):
  * ``trajectory`` with float64 1-D ``xs``, ``ys``, ``timestamp`` and ``add_point(time, x, y)``   (GPmap.py:12-23)
  * ``trajectories`` with ``pathdict``, ``add_trajectory(id, trajectory)``, ``kmeansclustering(k,
    treshold=1000)``, ``calc_mean_traj(keys)``, ``calc_distance(t1, t2)`` and the plot helpers      (GPmap.py:28-161)
  * ``check_if_valid_trajectory(traj, minimumtraveldistance=1)``                                  (GPmap.py:165-175)
  * ``readcsvfile(numoftrajstoread=0)`` filling the module-level ``trajs``                        (GPmap.py:163,178-204)
What is deliberately NOT kept: import-time execution (GPmap.py:212,220), the never-terminating
centroid re-draw loop (GPmap.py:39-53), the ZeroDivisionError on an empty cluster (GPmap.py:111),
and the cwd-relative hard-coded file name as the only option (GPmap.py:181).

What is new (the reference contains no GP code): ``fit_gp``, ``GPModel``,
``fit_gp_batched``, ``lml_sweep`` and ``trajectory.gp_inputs``.  They run on CUDA only, through the
C-ABI in ``include/gpmap_b200.h``; there is no CPU fallback.
"""
from __future__ import annotations

import csv
import ctypes as C
import random as rdm
import string

import numpy as np

from . import _native

__all__ = ["trajectory", "trajectories", "trajs", "check_if_valid_trajectory", "readcsvfile",
           "fit_gp", "GPModel", "fit_gp_batched", "lml_sweep", "make_theta", "export_raster", "optimize_gp",
           "clear_workspaces"]


# ------------------------------------------------------------------------------------------------
# device plumbing (PyTorch owns memory and streams; the math is in libgpmap_b200.so)
# ------------------------------------------------------------------------------------------------

def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("GPmap (B200-native) needs a CUDA device: there is no CPU fallback for the GP path")
    return torch


class _Stager:
    """Cached pinned staging buffers for host -> device copies (a fresh ``pin_memory()`` per call costs a
    ``cudaHostAlloc``).  A small ring of grow-only pinned buffers per device; a slot is reused only after the
    copy that last read it has completed (one event per slot)."""
    SLOTS = 4

    def __init__(self):
        self.slots = {}          # device index -> list of [pinned uint8 tensor | None, event | None]
        self.next = {}

    def to_device(self, arr, device):
        torch = _torch()
        dev = torch.device(device if device is not None else "cuda")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        ring = self.slots.setdefault(idx, [[None, None] for _ in range(self.SLOTS)])
        k = self.next.get(idx, 0)
        self.next[idx] = (k + 1) % self.SLOTS
        slot = ring[k]
        nbytes = arr.nbytes
        if slot[1] is not None:
            slot[1].synchronize()                      # the previous copy out of this slot has finished
        if slot[0] is None or slot[0].numel() < nbytes:
            slot[0] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
        host = slot[0][:nbytes].view(torch.float64).view(arr.shape)
        host.numpy()[...] = arr                        # one host memcpy into pinned memory
        out = host.to(torch.device("cuda", idx), non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(idx))
        return out


_stager = _Stager()


def _dev(a, device=None):
    """float64 contiguous CUDA tensor from a numpy array / sequence / tensor (zero-copy when already so).
    Host data goes through cached pinned staging buffers (one host memcpy + one asynchronous H2D copy)."""
    torch = _torch()
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64:
            t = t.double()
        if t.is_cuda:
            return t.contiguous()
        if t.is_pinned():
            return t.contiguous().to(device or "cuda", non_blocking=True)
        a = t.contiguous().numpy()
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if arr.size == 0:
        return torch.empty(arr.shape, dtype=torch.float64, device=device or "cuda")
    return _stager.to_device(arr, device)


# Scratch workspaces (the variance sweep's K*^T / W buffer, the batched fits' scratch) are cached per
# (device, stream, tag) and grow on demand, so repeated calls -- a sweep, a serving loop, a benchmark -- do not
# re-allocate gigabytes per call.  Work on different streams never shares a buffer.  ``workspace=`` arguments of
# the public functions override the cache with a caller-owned float64 CUDA tensor.
_workspaces = {}
PREDICT_WORKSPACE_BYTES = 4 << 30       # default cap of the variance workspace (queries are processed in chunks)


def _workspace(nbytes, device, tag, given=None):
    torch = _torch()
    if given is not None:
        if not (isinstance(given, torch.Tensor) and given.is_cuda and given.dtype == torch.float64 and
                given.is_contiguous()):
            raise ValueError("workspace must be a contiguous float64 CUDA tensor")
        return given
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream, tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() * 8 < nbytes:
        _workspaces.pop(key, None)
        ws = None
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        _workspaces[key] = ws
    return ws


def clear_workspaces():
    """Release every cached scratch workspace (device memory goes back to PyTorch's allocator)."""
    _workspaces.clear()


def _stream(t):
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _round_up(a, b):
    return (a + b - 1) // b * b


def make_theta(lengthscale, signal_var, noise_var, D):
    """theta = [l_1..l_D, signal_var, noise_var] from a scalar (isotropic) or per-dimension lengthscale."""
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (D,)).copy()
    return np.concatenate([ls, [float(signal_var), float(noise_var)]])


class GPModel:
    """Device-resident fitted model: X, theta, the Cholesky factor L (lower triangle of ``K``),
    the inverted diagonal blocks, alpha and the log marginal likelihood."""

    def __init__(self, X, theta, K, ws, alpha, lml, info):
        self.X, self.theta, self.K, self.ws, self.alpha, self.lml_dev, self.info = X, theta, K, ws, alpha, lml, info

    @property
    def N(self):
        return self.X.shape[0]

    @property
    def D(self):
        return self.X.shape[1]

    @property
    def L(self):
        """Lower Cholesky factor as a dense (N, N) tensor (copy; for inspection and tests)."""
        return self.K[:, : self.N].tril()

    @property
    def lml(self):
        return self.lml_dev.cpu().numpy()

    def check(self):
        """Raise numpy.linalg.LinAlgError if the factorisation met a non-positive pivot (synchronises)."""
        j = int(self.info.item())
        if j != 0:
            raise np.linalg.LinAlgError(f"covariance matrix is not positive definite: pivot {j} <= 0")
        return self

    # -- checkpoint / resume (the model handle is just device tensors) -------------
    def state_dict(self):
        """Everything needed to predict again, as CPU tensors / arrays (``torch.save``-able)."""
        return {"X": self.X.cpu(), "theta": np.asarray(self.theta), "K": self.K.cpu(), "ws": self.ws.cpu(),
                "alpha": self.alpha.cpu(), "lml": self.lml_dev.cpu(), "info": self.info.cpu(), "version": 1}

    @classmethod
    def from_state_dict(cls, state, device=None):
        torch = _torch()
        dev = device or "cuda"
        t = lambda a: a.to(dev) if isinstance(a, torch.Tensor) else torch.as_tensor(a).to(dev)   # noqa: E731
        return cls(t(state["X"]).contiguous(), np.asarray(state["theta"], dtype=np.float64), t(state["K"]).contiguous(),
                   t(state["ws"]).contiguous(), t(state["alpha"]).contiguous(), t(state["lml"]), t(state["info"]))

    def lml_grad(self):
        """d lml_r / d log(theta_j) as a numpy (R, D+2) array, j over [l_1..l_D, signal_var, noise_var]
        (R&W eq. 5.9; K^{-1} is formed on the tensor cores)."""
        torch = _torch()
        lib = _native.load()
        h = _native.handle(self.X.device.index or 0)
        R = self.alpha.shape[1]
        nbytes = int(lib.gpm_lml_grad_workspace_bytes(self.N))
        ws = torch.empty(nbytes // 8, dtype=torch.float64, device=self.X.device)
        grad = torch.empty((R, self.D + 2), dtype=torch.float64, device=self.X.device)
        rc = lib.gpm_lml_grad(h, _ptr(self.X), self.N, self.D, _native.theta_array(self.theta), _ptr(self.K),
                              self.K.stride(0), _ptr(self.ws), _ptr(self.alpha), R, _ptr(grad), _ptr(ws), nbytes,
                              _stream(self.X))
        _native.check(rc, "gpm_lml_grad")
        return grad.cpu().numpy()

    # -- prediction ---------------------------------------------------------------------------
    def _predict(self, Xs, grid, m0, m1, return_var, include_noise, workspace=None, out=None):
        torch = _torch()
        lib = _native.load()
        h = _native.handle(self.X.device.index or 0)
        M = m1 - m0
        R = self.alpha.shape[1]
        if out is not None:                  # caller-owned result buffers (e.g. slices of an all-gather buffer)
            mu, var = (out if return_var else (out, None))
            if mu.shape != (M, R) or not mu.is_contiguous() or (return_var and (var.shape != (M,) or not var.is_contiguous())):
                raise ValueError("out= buffers must be contiguous with shapes (M, R) and (M,)")
        else:
            mu = torch.empty((M, R), dtype=torch.float64, device=self.X.device)
            var = torch.empty((M,), dtype=torch.float64, device=self.X.device) if return_var else None
        flags = _native.PREDICT_MEAN | (_native.PREDICT_VAR if return_var else 0) | \
            (_native.PREDICT_ADD_NOISE if include_noise else 0)
        ws, nbytes = None, 0
        if return_var and M > 0:
            # the library sizes W for all M queries; cap it (queries are then processed in chunks of whole waves)
            nbytes = min(int(lib.gpm_predict_workspace_bytes(h, self.N, M)),
                         max(PREDICT_WORKSPACE_BYTES, int(lib.gpm_predict_workspace_bytes(h, self.N, 128 * 148))))
            ws = _workspace(nbytes, self.X.device, "predict", workspace)
            nbytes = ws.numel() * 8
        g = C.byref(grid) if grid is not None else None
        rc = lib.gpm_predict(h, _ptr(self.X), self.N, self.D, _native.theta_array(self.theta), _ptr(self.K),
                             self.K.stride(0), _ptr(self.ws), _ptr(self.alpha), R, _ptr(Xs), g, m0, m1,
                             _ptr(mu), _ptr(var), _ptr(ws), nbytes, flags, _stream(self.X))
        _native.check(rc, "gpm_predict")
        return (mu, var) if return_var else mu

    def predict(self, Xs, return_var=True, include_noise=False, workspace=None):
        """Posterior at arbitrary query points Xs (M, D): mu (M, R)[, var (M,)] as CUDA tensors.
        The variance (signal_var - ||L^-1 k*||^2) is clamped at 0: cancellation next to training points could
        otherwise return values of order -1e-16 signal_var."""
        Xs = _dev(Xs, self.X.device)
        if Xs.ndim != 2 or Xs.shape[1] != self.D:
            raise ValueError(f"Xs must be (M, {self.D})")
        return self._predict(Xs, None, 0, Xs.shape[0], return_var, include_noise, workspace)

    def predict_grid(self, bounds, shape, t=None, return_var=True, include_noise=False, points=None,
                     workspace=None, out=None):
        """Posterior on a regular grid (``indexing='xy'``, y outer).  bounds = (x0, x1, y0, y1),
        shape = (Gx, Gy).  Returns mu (Gy, Gx, R)[, var (Gy, Gx)]; with ``points=(m0, m1)`` only that
        flat range of grid points is evaluated and flat (m1-m0, R) / (m1-m0,) tensors are returned
        (this is how the grid is sharded across GPUs).  ``workspace``: caller-owned scratch (float64 CUDA tensor;
        default: a cached per-stream buffer of at most PREDICT_WORKSPACE_BYTES).  ``out=(mu, var)``: write the
        results of a ``points`` range into caller-owned flat buffers."""
        x0, x1, y0, y1 = [float(v) for v in bounds]
        Gx, Gy = int(shape[0]), int(shape[1])
        if self.D == 3 and t is None:
            raise ValueError("a D=3 model needs the query time t")
        grid = _native.GpmGrid(x0, x1, y0, y1, float(t) if t is not None else 0.0, Gx, Gy)
        m0, m1 = (0, Gx * Gy) if points is None else (int(points[0]), int(points[1]))
        out = self._predict(None, grid, m0, m1, return_var, include_noise, workspace, out)
        if points is not None:
            return out
        if return_var:
            return out[0].view(Gy, Gx, -1), out[1].view(Gy, Gx)
        return out.view(Gy, Gx, -1)


def export_raster(path, mu, var, bounds, shape):
    """Write posterior rasters over the reference's plot window (GPmap.py:126,155) as an .npz:
    mu (Gy, Gx, R), var (Gy, Gx), plus the axis vectors.  (matplotlib is not needed)"""
    import torch
    x0, x1, y0, y1 = [float(v) for v in bounds]
    Gx, Gy = int(shape[0]), int(shape[1])
    to_np = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)  # noqa: E731
    np.savez_compressed(path, mu=to_np(mu).reshape(Gy, Gx, -1), var=to_np(var).reshape(Gy, Gx),
                        x=np.linspace(x0, x1, Gx), y=np.linspace(y0, y1, Gy), bounds=np.array([x0, x1, y0, y1]))
    return path


def fit_gp(X, Y, lengthscale=None, signal_var=1.0, noise_var=1e-2, theta=None, check=True, lml=True):
    """Fit one GP: K = k(X,X) + noise_var I, L = chol(K), alpha = K^{-1} Y, LML.  (R&W Alg. 2.1.)

    X (N, D) with D in {2, 3} (columns xs, ys[, timestamp]); Y (N,) or (N, R).  numpy arrays are
    copied to the current CUDA device; CUDA float64 tensors are used in place.  Returns a GPModel.
    """
    torch = _torch()
    lib = _native.load()
    X = _dev(X)
    Y = _dev(Y, X.device)
    if Y.ndim == 1:
        Y = Y[:, None].contiguous()
    N, D = X.shape
    if D not in (2, 3):
        raise ValueError("X must have 2 or 3 columns (xs, ys[, timestamp])")
    if Y.shape[0] != N:
        raise ValueError("X and Y disagree on N")
    R = Y.shape[1]
    if theta is None:
        if lengthscale is None:
            raise ValueError("give lengthscale= or theta=")
        theta = make_theta(lengthscale, signal_var, noise_var, D)
    theta = np.asarray(theta, dtype=np.float64)
    if theta.shape != (D + 2,):
        raise ValueError(f"theta must have {D + 2} entries")
    h = _native.handle(X.device.index or 0)
    st = _stream(X)
    ld = _round_up(N, 16)
    K = torch.empty((N, ld), dtype=torch.float64, device=X.device)
    ws = torch.empty(int(lib.gpm_potrf_workspace_bytes(N)) // 8, dtype=torch.float64, device=X.device)
    info = torch.zeros(1, dtype=torch.int32, device=X.device)
    alpha = torch.empty((N, R), dtype=torch.float64, device=X.device)
    lml_dev = torch.full((R,), float("nan"), dtype=torch.float64, device=X.device)    # stays NaN with lml=False
    th = _native.theta_array(theta)
    # one call: covariance -> Cholesky (forward substitution riding along) -> backward substitution -> LML
    _native.check(lib.gpm_fit(h, _ptr(X), N, D, th, _ptr(Y), R, _ptr(K), ld, _ptr(ws), _ptr(alpha),
                              _ptr(lml_dev) if lml else C.c_void_p(0), _ptr(info), st), "gpm_fit")
    model = GPModel(X, theta, K, ws, alpha, lml_dev, info)
    if check:
        model.check()
    return model


def fit_gp_batched(Xb, Yb, lengthscale=None, signal_var=1.0, noise_var=1e-2, theta=None, check=True, workspace=None,
                   out=None):
    """B independent equal-length paths: Xb (B, N, D), Yb (B, N, R) -> alpha (B, N, R), lml (B, R) (CUDA).

    Paths of up to 112 samples (the reference resamples to 33, GPmap.py:189) are fitted one CTA per path entirely in
    shared memory; up to 1024 samples one CTA per path with tensor-core tiles (two paths in flight per SM); longer
    ones go through the tiled batched pipeline (B <= 65535 there).  ``workspace``: caller-owned scratch;
    ``out=(alpha, lml)``: caller-owned result buffers."""
    torch = _torch()
    lib = _native.load()
    Xb = _dev(Xb)
    Yb = _dev(Yb, Xb.device)
    if Yb.ndim == 2:
        Yb = Yb[:, :, None].contiguous()
    B, N, D = Xb.shape
    R = Yb.shape[2]
    if theta is None:
        theta = make_theta(lengthscale, signal_var, noise_var, D)
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))     # (D+2,) shared or (B, D+2) per path
    if theta.shape not in ((D + 2,), (B, D + 2)):
        raise ValueError(f"theta must have shape ({D + 2},) or ({B}, {D + 2})")
    stride = D + 2 if theta.ndim == 2 else 0
    h = _native.handle(Xb.device.index or 0)
    if out is not None:                      # caller-owned results (e.g. this rank's slices of an all-gather buffer)
        alpha, lml = out
        if alpha.shape != (B, N, R) or lml.shape != (B, R) or not alpha.is_contiguous() or not lml.is_contiguous():
            raise ValueError("out= must be contiguous (B, N, R) and (B, R) float64 CUDA tensors")
    else:
        alpha = torch.empty((B, N, R), dtype=torch.float64, device=Xb.device)
        lml = torch.empty((B, R), dtype=torch.float64, device=Xb.device)
    info = torch.zeros(B, dtype=torch.int32, device=Xb.device)
    ws = _workspace(int(lib.gpm_fit_batched_workspace_bytes(h, B, N)), Xb.device, "fit_batched", workspace)
    th_arr = _native.theta_array(theta.ravel())
    rc = lib.gpm_fit_batched(h, _ptr(Xb), _ptr(Yb), B, N, D, R, th_arr, stride,
                             _ptr(alpha), _ptr(lml), _ptr(info), _ptr(ws), _stream(Xb))
    _native.check(rc, "gpm_fit_batched")
    if stride:
        torch.cuda.current_stream(Xb.device).synchronize()      # the host theta array is read by an async copy
    if check:
        bad = torch.nonzero(info)
        if bad.numel():
            b = int(bad[0])
            raise np.linalg.LinAlgError(f"path {b}: covariance not positive definite (pivot {int(info[b])})")
    return alpha, lml


def lml_sweep(X, Y, thetas, indices=None):
    """Log marginal likelihood at hyper-parameter points thetas (S, D+2) -> lml (S, R) numpy.
    ``indices`` restricts the sweep to a subset (used to shard the sweep across ranks)."""
    X = _dev(X)
    Y = _dev(Y, X.device)
    thetas = np.asarray(thetas, dtype=np.float64)
    idx = range(thetas.shape[0]) if indices is None else list(indices)
    out = []
    for s in idx:
        m = fit_gp(X, Y, theta=thetas[s], check=False)
        out.append(m.lml_dev)
        del m
    torch = _torch()
    if not out:
        return np.empty((0, 1 if Y.ndim == 1 else Y.shape[1]))
    return torch.stack(out).cpu().numpy()


def optimize_gp(X, Y, theta0, maxiter=50, bounds=None, tie_xy=True):
    """Hyper-parameter learning: maximise sum_r lml_r over log(theta) with L-BFGS-B (scipy on the host,
    fit + exact gradient on the GPU).  ``tie_xy`` keeps one lengthscale for x and y (isotropic in space),
    as the sweep of BASELINE config 5 does.  Returns (theta, model, scipy result)."""
    from scipy.optimize import minimize
    X = _dev(X)
    Y = _dev(Y, X.device)
    D = X.shape[1]
    theta0 = np.asarray(theta0, dtype=np.float64)

    def unpack(u):
        th = np.exp(u)
        if tie_xy:
            th = np.concatenate([[th[0]], th])          # u = [l_xy, (l_t), sf2, sn2]
        return th

    def fun(u):
        th = unpack(u)
        try:
            m = fit_gp(X, Y, theta=th, check=True)
        except np.linalg.LinAlgError:
            return 1e300, np.zeros_like(u)
        g = m.lml_grad().sum(axis=0)
        if tie_xy:
            g = np.concatenate([[g[0] + g[1]], g[2:]])
        return -float(m.lml.sum()), -g

    u0 = np.log(theta0[1:] if tie_xy else theta0)
    res = minimize(fun, u0, jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": maxiter})
    theta = unpack(res.x)
    return theta, fit_gp(X, Y, theta=theta), res


# ------------------------------------------------------------------------------------------------
# the reference's data model
# ------------------------------------------------------------------------------------------------

class trajectory:
    """One path: three growing float64 arrays (GPmap.py:12-23)."""

    def __init__(self):
        self.xs = np.array([], dtype=float)
        self.ys = np.array([], dtype=float)
        self.timestamp = np.array([], dtype=float)

    def add_point(self, time, x, y):
        self.xs = np.append(self.xs, x)
        self.ys = np.append(self.ys, y)
        self.timestamp = np.append(self.timestamp, time)

    def get_trajectory(self):
        """The reference leaves this as a stub printing "not implemented" (GPmap.py:25-26); here it
        returns the samples as an (n, 3) array of (timestamp, x, y)."""
        return np.stack([self.timestamp, self.xs, self.ys], axis=1)

    def gp_inputs(self, use_time=False):
        """GP input matrix X (n, D): columns xs, ys[, timestamp]."""
        cols = [self.xs, self.ys] + ([self.timestamp] if use_time else [])
        return np.ascontiguousarray(np.stack(cols, axis=1))

    def __len__(self):
        return len(self.xs)


class trajectories:
    """id -> trajectory map with the reference's clustering helpers (GPmap.py:28-161)."""

    def __init__(self):
        self.pathdict = {}

    def add_trajectory(self, id, trajectory):
        self.pathdict[id] = trajectory

    # -- distances and means (host; same arithmetic order as the reference) --------------------
    def calc_distance(self, traj1, traj2):
        """Sum over corresponding samples of the point distance (GPmap.py:114-121); assumes equal length."""
        dx = np.asarray(traj1.xs, dtype=float) - np.asarray(traj2.xs, dtype=float)
        dy = np.asarray(traj1.ys, dtype=float) - np.asarray(traj2.ys, dtype=float)
        d = np.sqrt(dx * dx + dy * dy)
        return float(np.cumsum(d)[-1]) if len(d) else 0.0      # cumsum = the reference's sequential sum

    def calc_mean_traj(self, traj):
        """Point-wise mean trajectory of the listed path ids (GPmap.py:95-112)."""
        if len(traj) == 0:
            raise ValueError("calc_mean_traj needs at least one member (the reference divides by zero here)")
        members = [self.pathdict[key] for key in traj]
        sx = np.zeros_like(members[0].xs)
        sy = np.zeros_like(members[0].ys)
        stime = np.zeros_like(members[0].timestamp)
        for m in members:                       # member-order accumulation, as the reference sums
            sx = sx + m.xs
            sy = sy + m.ys
            stime = stime + m.timestamp
        out = trajectory()
        out.xs, out.ys, out.timestamp = sx / len(members), sy / len(members), stime / len(members)
        return out

    def packed(self, keys=None, use_time=False):
        """Stack equal-length paths into X (B, n, D) in key order (the batched-fit input)."""
        keys = list(self.pathdict.keys()) if keys is None else list(keys)
        return np.stack([self.pathdict[k].gp_inputs(use_time) for k in keys]), keys

    # -- k-means (GPmap.py:36-93): assignment + centroid update + convergence test run on the GPU ----------
    def _upload_paths(self, keys):
        """One host -> device copy of all paths: (3, P, n) = xs, ys, timestamp planes, path-major; the
        sample-major (n, P) copies the assignment kernel reads are made on the device."""
        host = np.stack([np.stack([self.pathdict[k].xs for k in keys]),
                         np.stack([self.pathdict[k].ys for k in keys]),
                         np.stack([self.pathdict[k].timestamp for k in keys])])
        dev = _dev(host)
        pxT, pyT = dev[0].t().contiguous(), dev[1].t().contiguous()
        return dev, pxT, pyT

    def _assign(self, keys, centroids):
        """One assignment step (GPmap.py:72-80): returns (assign (P,), dist (P, k)) as numpy arrays."""
        torch = _torch()
        lib = _native.load()
        dev, pxT, pyT = self._upload_paths(keys)
        cx = _dev(np.stack([c.xs for c in centroids]), dev.device)
        cy = _dev(np.stack([c.ys for c in centroids]), dev.device)
        _, P, n = dev.shape
        k = cx.shape[0]
        dist = torch.empty((P, k), dtype=torch.float64, device=dev.device)
        assign = torch.empty((P,), dtype=torch.int32, device=dev.device)
        h = _native.handle(dev.device.index or 0)
        rc = lib.gpm_kmeans_assign(h, _ptr(pxT), _ptr(pyT), P, n, _ptr(cx), _ptr(cy), k, _ptr(dist), _ptr(assign),
                                   _stream(dev))
        _native.check(rc, "gpm_kmeans_assign")
        return assign.cpu().numpy(), dist.cpu().numpy()

    def kmeansclustering(self, k, treshold=1000, plot=False, max_iter=1000, seed=None, init=None):
        """Lloyd iterations with the reference's trajectory distance, stopping when the summed centroid
        shift drops below 5 (GPmap.py:90).  Returns {centroid_key: [path ids]}.

        The paths are uploaded once; assignment, centroid means (members summed in path order, bit-exact with
        ``calc_mean_traj``) and the convergence sum run on the device, several iterations per host round trip;
        only the final assignment and centroids come back.

        Differences from the reference, all bug fixes: initial centroids closer than ``treshold`` are
        actually re-drawn (bounded retries; the reference's loop never terminates, GPmap.py:39-53); an
        empty cluster keeps its previous centroid (the reference raises ZeroDivisionError, GPmap.py:111);
        plotting is opt-in.  ``init`` (k path ids) fixes the initial centroids instead of drawing them."""
        torch = _torch()
        lib = _native.load()
        rng = rdm.Random(seed) if seed is not None else rdm
        keys = list(self.pathdict.keys())
        if k > len(keys):
            raise ValueError("more clusters than trajectories")
        chosen = list(init) if init is not None else rng.sample(keys, k)
        if len(chosen) != k:
            raise ValueError("init must name k paths")
        for _ in range(0 if init is not None else 100):
            close = [(i, j) for i in range(k) for j in range(i + 1, k)
                     if self.calc_distance(self.pathdict[chosen[i]], self.pathdict[chosen[j]]) < treshold]
            free = [key for key in keys if key not in chosen]
            if not close or not free:
                break
            chosen[close[0][1]] = rng.choice(free)
        names = []
        while len(names) < k:
            name = "".join(rng.choices(string.ascii_uppercase + string.digits, k=5))
            if name not in names:
                names.append(name)
        dev, pxT, pyT = self._upload_paths(keys)                       # the one H2D copy of the call
        _, P, n = dev.shape
        index = {key: i for i, key in enumerate(keys)}
        sel = torch.tensor([index[key] for key in chosen], dtype=torch.long, device=dev.device)
        cents = dev[:, sel, :].contiguous()                           # (3, k, n): initial centroids = chosen paths
        assign = torch.zeros((P,), dtype=torch.int32, device=dev.device)
        ws = torch.empty(int(lib.gpm_kmeans_workspace_bytes(P, n, k)) // 8, dtype=torch.float64, device=dev.device)
        h = _native.handle(dev.device.index or 0)
        st = _stream(dev)
        batch = 8                                                      # iterations enqueued per host round trip
        iters, conv, shift = C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        done = 0
        while done < max_iter:
            todo = min(batch, max_iter - done)
            todo += todo & 1                                           # even counts keep the ping-pong parity simple
            rc = lib.gpm_kmeans_lloyd(h, _ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(pxT), _ptr(pyT), P, n, k,
                                      _ptr(cents), _ptr(assign), 5.0, todo, 1 if done == 0 else 0, _ptr(ws), st)
            _native.check(rc, "gpm_kmeans_lloyd")
            _native.check(lib.gpm_kmeans_state(h, _ptr(ws), n, k, C.byref(iters), C.byref(conv), C.byref(shift), st),
                          "gpm_kmeans_state")
            done += todo
            if conv.value:
                break
        final = ws[: 3 * k * n].view(3, k, n) if (iters.value & 1) else cents
        final = final.cpu().numpy()
        assign_h = assign.cpu().numpy()
        self.kmeans_iterations, self.kmeans_shift = int(iters.value), float(shift.value)
        clusters = {name: [] for name in names}
        for key, a in zip(keys, assign_h):
            clusters[names[int(a)]].append(key)
        self.centroids = {}
        for i, name in enumerate(names):
            c = trajectory()
            c.xs, c.ys, c.timestamp = final[0, i].copy(), final[1, i].copy(), final[2, i].copy()
            self.centroids[name] = c
        if plot:
            self.plotclusters(clusters)
        return clusters

    # -- plotting (GPmap.py:125-161); matplotlib is optional --------------------------------------
    @staticmethod
    def _plt():
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:        # pragma: no cover - matplotlib is not in this image
            raise RuntimeError("plotting needs matplotlib, which is not installed") from e
        return plt

    def plotclusters(self, clusters):
        plt = self._plt()
        plt.axis([-50000, 50000.0, -50000.0, 50000.0])
        for members in clusters.values():
            colour = "#%06X" % rdm.randint(0, 0xFFFFFF)
            for key in members:
                plt.plot(self.pathdict[key].xs, self.pathdict[key].ys, colour)
        plt.show()

    def plotwx(self, x):
        plt = self._plt()
        for id in x:
            plt.plot(x[id].xs, x[id].ys, "*")
        self.plot()

    def plot(self):
        plt = self._plt()
        plt.axis([-50000, 50000.0, -50000.0, 50000.0])
        for id in self.pathdict:
            plt.plot(self.pathdict[id].xs, self.pathdict[id].ys)
        plt.show()

    # -- GP fits over the stored paths -------------------------------------------------------------
    def fit_gp_batched(self, targets, lengthscale=None, signal_var=1.0, noise_var=1e-2, theta=None,
                       use_time=True, keys=None):
        """Fit every stored path independently.  targets: {id: (n, R) array} or (B, n, R) in key order."""
        Xb, keys = self.packed(keys, use_time)
        Yb = np.stack([np.asarray(targets[k], dtype=float) for k in keys]) if isinstance(targets, dict) \
            else np.asarray(targets, dtype=float)
        alpha, lml = fit_gp_batched(Xb, Yb, lengthscale, signal_var, noise_var, theta)
        return alpha, lml, keys


trajs = trajectories()


def check_if_valid_trajectory(traj, minimumtraveldistance=1):
    """The reference's ingest filter (GPmap.py:165-175): sum over i<j of (|x_j|-|x_i|) + (|y_j|-|y_i|)
    compared with the threshold, evaluated in closed form sum_m (2m-(n-1)) (|x_m|+|y_m|)."""
    n = len(traj.xs)
    w = 2.0 * np.arange(n) - (n - 1)
    total = float(np.sum(w * (np.abs(traj.xs) + np.abs(traj.ys))))
    return not (total < minimumtraveldistance)


def readcsvfile(numoftrajstoread=0, filename="testfile.csv", target=None, samples=33, threshold=1000):
    """Parse the reference's CSV format into ``trajs`` (GPmap.py:178-204): a header row carries the id in
    column 1, data rows are (time, _, x, y) with integer coordinates, a row starting '###' ends a path;
    only paths with exactly ``samples`` points that pass the validity filter are kept."""
    dest = trajs if target is None else target
    kept = 0
    current, current_id, in_header = None, 0, True
    with open(filename, newline="") as fh:
        for row in csv.reader(fh, delimiter=","):
            if not row:
                continue
            if row[0] == "###":
                if current is not None and len(current.timestamp) == samples and \
                        check_if_valid_trajectory(current, threshold):
                    dest.add_trajectory(current_id, current)
                    kept += 1
                if numoftrajstoread != 0 and kept >= numoftrajstoread:
                    break
                current, in_header = trajectory(), True
            elif in_header:
                current_id, current, in_header = row[1], trajectory(), False
            else:
                current.add_point(float(row[0]), int(row[2]), int(row[3]))
    return dest
