"""CPU oracle for the GP path-modelling hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product (``gaussianprocesspathmodelling_b200``) never does.

PARITY UNPINNED BY THE REFERENCE.  ``/root/reference/GPmap.py`` (219 lines) contains no
Gaussian-process code at all: no covariance, no Cholesky, no likelihood, no posterior (SURVEY.md
section 0).  There is therefore no reference implementation, golden vector or fixture for this path.
What this file restates is the textbook algorithm BASELINE.json's ``north_star`` describes --
Rasmussen & Williams, *Gaussian Processes for Machine Learning* (2006), Algorithm 2.1, with a
squared-exponential (RBF/ARD) kernel -- written with the two numeric libraries the reference
imports (numpy ``GPmap.py:1``; scipy.spatial.distance ``GPmap.py:10``) and consuming inputs in the
reference's data-model layout (``trajectory.xs/.ys/.timestamp`` float64 1-D arrays,
``GPmap.py:15-23``).  The oracle itself is pinned by this repo's tests against
  * scikit-learn's GaussianProcessRegressor (independent implementation),
  * mpmath at 50 digits for N <= 48,
  * hand-computed N=1 / N=2 posteriors and the iid-noise limit of the LML,
see ``tests/test_oracle.py``.

Conventions (shared with the CUDA path, see ``include/gpmap_b200.h``):
  X      (N, D) float64, D in {2, 3}; columns = xs, ys[, timestamp]
  Y      (N, R) float64 targets
  theta  (D + 2,) float64 = [l_1 .. l_D, signal_var, noise_var]
  k(x, z) = signal_var * exp(-0.5 * sum_d ((x_d - z_d) / l_d)^2)
  K      = k(X, X) + noise_var * I
  L      lower Cholesky factor, K = L L^T
  alpha  = K^{-1} Y;   lml_r = -0.5 Y_r^T alpha_r - sum_i log L_ii - (N/2) log(2 pi)
  mean   = k(Xs, X) alpha;  var = signal_var - sum_i (L^{-1} k(X, Xs))_i^2   (latent-f variance)
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular
from scipy.spatial.distance import cdist

LOG_2PI = math.log(2.0 * math.pi)


def split_theta(theta, D):
    """theta = [l_1..l_D, signal_var, noise_var] -> (lengthscales (D,), signal_var, noise_var)."""
    theta = np.asarray(theta, dtype=np.float64)
    if theta.shape != (D + 2,):
        raise ValueError(f"theta must have shape ({D + 2},), got {theta.shape}")
    return theta[:D].copy(), float(theta[D]), float(theta[D + 1])


def make_theta(lengthscale, signal_var, noise_var, D):
    """Build theta from a scalar (isotropic) or per-dimension lengthscale."""
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (D,)).copy()
    return np.concatenate([ls, [float(signal_var), float(noise_var)]])


def cross_cov(X, Z, theta):
    """k(X, Z): (N, M).  R&W eq. (2.16) squared-exponential, ARD lengthscales.  No noise term."""
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    ls, sf2, _ = split_theta(theta, X.shape[1])
    d2 = cdist(X / ls, Z / ls, "sqeuclidean")
    return sf2 * np.exp(-0.5 * d2)


def cov(X, theta):
    """K(X, X) + noise_var * I: (N, N).  R&W Alg. 2.1 input to line 2."""
    X = np.asarray(X, dtype=np.float64)
    _, _, sn2 = split_theta(theta, X.shape[1])
    K = cross_cov(X, X, theta)
    K[np.diag_indices_from(K)] += sn2
    return K


def fit(X, Y, theta):
    """R&W Alg. 2.1 lines 2-3 and 7.  Returns dict(X, theta, L, alpha, lml)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    N = X.shape[0]
    K = cov(X, theta)
    L = cholesky(K, lower=True)                                   # line 2
    z = solve_triangular(L, Y, lower=True)                        # line 3 (inner solve)
    alpha = solve_triangular(L, z, lower=True, trans="T")         # line 3 (outer solve)
    lml = (-0.5 * np.einsum("nr,nr->r", Y, alpha)
           - np.log(np.diag(L)).sum() - 0.5 * N * LOG_2PI)        # line 7
    return {"X": X, "theta": np.asarray(theta, dtype=np.float64), "L": L, "alpha": alpha, "lml": lml}


def predict(model, Xs, return_var=True, include_noise=False, tile=8192):
    """R&W Alg. 2.1 lines 4-6 at query points Xs (M, D), tiled over M to bound memory."""
    Xs = np.asarray(Xs, dtype=np.float64)
    X, theta, L, alpha = model["X"], model["theta"], model["L"], model["alpha"]
    _, sf2, sn2 = split_theta(theta, X.shape[1])
    M = Xs.shape[0]
    mu = np.empty((M, alpha.shape[1]))
    var = np.empty(M) if return_var else None
    for m0 in range(0, M, tile):
        m1 = min(M, m0 + tile)
        Ks = cross_cov(X, Xs[m0:m1], theta)                       # (N, Mt)
        mu[m0:m1] = Ks.T @ alpha                                  # line 4
        if return_var:
            V = solve_triangular(L, Ks, lower=True)               # line 5
            var[m0:m1] = sf2 - np.einsum("nm,nm->m", V, V)        # line 6
            if include_noise:
                var[m0:m1] += sn2
    return (mu, var) if return_var else mu


def grid_points(bounds, shape, t=None):
    """Regular query grid, ``indexing='xy'``, row-major with y outer: point m = gy * Gx + gx.

    bounds = (x0, x1, y0, y1); shape = (Gx, Gy).  Coordinates are np.linspace values (endpoints
    included).  If ``t`` is given a constant third column is appended (D = 3 models).
    """
    x0, x1, y0, y1 = bounds
    Gx, Gy = shape
    gx = np.linspace(x0, x1, Gx)
    gy = np.linspace(y0, y1, Gy)
    xx, yy = np.meshgrid(gx, gy, indexing="xy")
    cols = [xx.ravel(), yy.ravel()]
    if t is not None:
        cols.append(np.full(Gx * Gy, float(t)))
    return np.stack(cols, axis=1)


def predict_grid(model, bounds, shape, t=None, return_var=True, include_noise=False):
    """Posterior on the regular grid: mu (Gy, Gx, R), var (Gy, Gx)."""
    Gx, Gy = shape
    out = predict(model, grid_points(bounds, shape, t), return_var=return_var, include_noise=include_noise)
    if return_var:
        mu, var = out
        return mu.reshape(Gy, Gx, -1), var.reshape(Gy, Gx)
    return out.reshape(Gy, Gx, -1)


def fit_batched(Xb, Yb, theta):
    """Independent fits of B equal-length paths (the reference assumes equal length, GPmap.py:96,117).

    Xb (B, N, D), Yb (B, N, R), theta (D+2,) shared or (B, D+2) per path.
    Returns alpha (B, N, R), lml (B, R).  Plain Python loop, as the reference's style would do.
    """
    Xb = np.asarray(Xb, dtype=np.float64)
    Yb = np.asarray(Yb, dtype=np.float64)
    B, N, D = Xb.shape
    theta = np.asarray(theta, dtype=np.float64)
    alpha = np.empty_like(Yb)
    lml = np.empty((B, Yb.shape[2]))
    for b in range(B):
        th = theta[b] if theta.ndim == 2 else theta
        m = fit(Xb[b], Yb[b], th)
        alpha[b] = m["alpha"]
        lml[b] = m["lml"]
    return alpha, lml


def lml_sweep(X, Y, thetas):
    """LML at S hyper-parameter points: thetas (S, D+2) -> lml (S, R)."""
    thetas = np.asarray(thetas, dtype=np.float64)
    return np.stack([fit(X, Y, th)["lml"] for th in thetas])


def lml_grad(X, Y, theta):
    """Gradient of the log marginal likelihood w.r.t. the LOG hyper-parameters (R&W eq. 5.9):

        d lml_r / d log(theta_j) = 0.5 * alpha_r^T dK_j alpha_r - 0.5 * tr(K^{-1} dK_j),
        dK_j = dK / d log(theta_j):   l_d -> K_f * ((x_d - z_d) / l_d)^2,   signal_var -> K_f,
                                      noise_var -> noise_var * I.
    Returns (R, D + 2)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    N, D = X.shape
    ls, sf2, sn2 = split_theta(theta, D)
    m = fit(X, Y, theta)
    Linv = solve_triangular(m["L"], np.eye(N), lower=True)
    Kinv = Linv.T @ Linv
    Kf = cross_cov(X, X, theta)
    dKs = []
    for d in range(D):
        diff = (X[:, d][:, None] - X[:, d][None, :]) / ls[d]
        dKs.append(Kf * diff * diff)
    dKs.append(Kf)
    dKs.append(sn2 * np.eye(N))
    a = m["alpha"]
    g = np.empty((Y.shape[1], D + 2))
    for j, dK in enumerate(dKs):
        tr = np.sum(Kinv * dK)
        g[:, j] = 0.5 * np.einsum("nr,nm,mr->r", a, dK, a) - 0.5 * tr
    return g


# ----------------------------------------------------------------------------------------------
# Restatement of the code the reference DOES contain (SURVEY.md section 8f "next" rows): the
# trajectory distance, centroid mean and ingest filter.  These are pinned against the reference
# itself (imported with a matplotlib stub) by tests/golden/make_reference_golden.py.
# ----------------------------------------------------------------------------------------------

def calc_distance(x1, y1, x2, y2):
    """Sum over corresponding samples of the Euclidean distance (GPmap.py:114-121)."""
    # the reference accumulates sequentially in a Python float; np.linalg.norm of a 1x2 matrix
    # is sqrt(dx*dx + dy*dy)
    s = 0.0
    for i in range(len(x1)):
        dx = x1[i] - x2[i]
        dy = y1[i] - y2[i]
        s += float(np.sqrt(dx * dx + dy * dy))
    return s


def calc_mean_traj(xs, ys, ts):
    """Point-wise mean of member trajectories (GPmap.py:95-112). xs, ys, ts: (members, n)."""
    n_members = xs.shape[0]
    mx = np.zeros(xs.shape[1]); my = np.zeros(xs.shape[1]); mt = np.zeros(xs.shape[1])
    for v in range(n_members):          # sequential member order, as the reference sums
        mx = mx + xs[v]; my = my + ys[v]; mt = mt + ts[v]
    return mx / n_members, my / n_members, mt / n_members


def travel_sum(xs, ys):
    """The double loop of check_if_valid_trajectory (GPmap.py:165-175), in closed form.

    sum_{i<j} (|x_j| - |x_i|) + (|y_j| - |y_i|) = sum_m (2m - (n-1)) (|x_m| + |y_m|).
    Exact for the integer-valued coordinates the reference ingests (GPmap.py:199).
    """
    n = len(xs)
    w = 2.0 * np.arange(n) - (n - 1)
    return float(np.sum(w * (np.abs(xs) + np.abs(ys))))


def lloyd(xs, ys, ts, init_idx, threshold=5.0, max_iter=1000):
    """The Lloyd loop of trajectories.kmeansclustering (GPmap.py:65-93) from fixed initial centroids.

    xs, ys, ts: (P, n).  init_idx: indices of the k paths copied as initial centroids (GPmap.py:58-60).
    Every path joins the first centroid with the smallest calc_distance (strict '<', GPmap.py:72-80), centroids
    become calc_mean_traj of their members in path order (GPmap.py:83-84), and the loop stops when the summed
    centroid shift is below ``threshold`` (GPmap.py:87-90).  Returns (assign (P,), centroids (3, k, n), iterations).
    An empty cluster keeps its centroid (the reference divides by zero there, GPmap.py:111).
    """
    P, n = xs.shape
    k = len(init_idx)
    cx = xs[list(init_idx)].copy(); cy = ys[list(init_idx)].copy(); ct = ts[list(init_idx)].copy()
    assign = np.zeros(P, dtype=np.int64)
    for it in range(1, max_iter + 1):
        for p in range(P):
            best, best_c = None, 0
            for c in range(k):
                d = calc_distance(cx[c], cy[c], xs[p], ys[p])
                if best is None or d < best:
                    best, best_c = d, c
            assign[p] = best_c
        nx, ny, nt = cx.copy(), cy.copy(), ct.copy()
        for c in range(k):
            mem = np.nonzero(assign == c)[0]
            if len(mem):
                nx[c], ny[c], nt[c] = calc_mean_traj(xs[mem], ys[mem], ts[mem])
        tot = 0
        for c in range(k):
            tot += calc_distance(nx[c], ny[c], cx[c], cy[c])
        cx, cy, ct = nx, ny, nt
        if tot < threshold:
            return assign, np.stack([cx, cy, ct]), it
    return assign, np.stack([cx, cy, ct]), max_iter


# ----------------------------------------------------------------------------------------------
# Timed variants for bench.py's CPU baseline: the same calls as fit() / predict(), with wall-clock per step
# (cov / cholesky / solve+LML / mean / variance reported separately).
# ----------------------------------------------------------------------------------------------

def fit_phases(X, Y, theta):
    """fit() with per-step wall times: returns (model, {"cov_s", "cholesky_s", "solve_lml_s"})."""
    import time
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    N = X.shape[0]
    t0 = time.perf_counter()
    K = cov(X, theta)
    t1 = time.perf_counter()
    L = cholesky(K, lower=True)
    t2 = time.perf_counter()
    z = solve_triangular(L, Y, lower=True)
    alpha = solve_triangular(L, z, lower=True, trans="T")
    lml = -0.5 * np.einsum("nr,nr->r", Y, alpha) - np.log(np.diag(L)).sum() - 0.5 * N * LOG_2PI
    t3 = time.perf_counter()
    model = {"X": X, "theta": np.asarray(theta, dtype=np.float64), "L": L, "alpha": alpha, "lml": lml}
    return model, {"cov_s": t1 - t0, "cholesky_s": t2 - t1, "solve_lml_s": t3 - t2}


def predict_phases(model, Xs, tile=8192):
    """predict() with per-step wall times: returns (mu, var, {"cross_cov_s", "mean_s", "var_s"})."""
    import time
    Xs = np.asarray(Xs, dtype=np.float64)
    X, theta, L, alpha = model["X"], model["theta"], model["L"], model["alpha"]
    _, sf2, _ = split_theta(theta, X.shape[1])
    M = Xs.shape[0]
    mu = np.empty((M, alpha.shape[1]))
    var = np.empty(M)
    tm = {"cross_cov_s": 0.0, "mean_s": 0.0, "var_s": 0.0}
    for m0 in range(0, M, tile):
        m1 = min(M, m0 + tile)
        t0 = time.perf_counter()
        Ks = cross_cov(X, Xs[m0:m1], theta)
        t1 = time.perf_counter()
        mu[m0:m1] = Ks.T @ alpha
        t2 = time.perf_counter()
        V = solve_triangular(L, Ks, lower=True)
        var[m0:m1] = sf2 - np.einsum("nm,nm->m", V, V)
        t3 = time.perf_counter()
        tm["cross_cov_s"] += t1 - t0; tm["mean_s"] += t2 - t1; tm["var_s"] += t3 - t2
    return mu, var, tm
