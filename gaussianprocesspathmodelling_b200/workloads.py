"""Seeded synthetic workloads for the five BASELINE.json configs.

numpy only; consumed byte-identically by the CUDA path, the oracle, the tests and bench.py.
Paths live in the reference's coordinate box (+-5e4, ``GPmap.py:126``), coordinates are rounded to
integers as the reference's ingest does (``int(row[2])``, ``GPmap.py:199``) and stored float64
(``GPmap.py:16-18``); timestamps are float seconds.
"""
from __future__ import annotations

import numpy as np

BOX = (-5.0e4, 5.0e4, -5.0e4, 5.0e4)          # x0, x1, y0, y1  (GPmap.py:126)
LENGTHSCALE_XY = 8000.0
LENGTHSCALE_T = 4.0
SIGNAL_VAR = 1.0
NOISE_VAR = 1.0e-2


def make_path(N, rng):
    """One smooth noisy 2-D path: returns xs, ys, timestamp (float64, length N)."""
    t = np.linspace(0.0, 1.0, N)
    f1, f2 = rng.uniform(0.5, 2.0, size=2)
    p1, p2 = rng.uniform(0.0, 2.0 * np.pi, size=2)
    x = 4.0e4 * np.cos(2.0 * np.pi * f1 * t + p1) + 500.0 * rng.standard_normal(N)
    y = 4.0e4 * np.sin(2.0 * np.pi * f2 * t + p2) + 500.0 * rng.standard_normal(N)
    xs = np.rint(x).astype(np.float64)
    ys = np.rint(y).astype(np.float64)
    ts = 0.5 * np.arange(N, dtype=np.float64)
    return xs, ys, ts


def path_targets(xs, ys, ts, R=2):
    """Targets: finite-difference velocities (dx/dt, dy/dt), standardised to unit variance."""
    vx = np.gradient(xs, ts)
    vy = np.gradient(ys, ts)
    Y = np.stack([vx, vy], axis=1)
    Y = (Y - Y.mean(axis=0)) / Y.std(axis=0)
    return np.ascontiguousarray(Y[:, :R])


def default_theta(D):
    ls = [LENGTHSCALE_XY, LENGTHSCALE_XY] + ([LENGTHSCALE_T] if D == 3 else [])
    return np.array(ls + [SIGNAL_VAR, NOISE_VAR], dtype=np.float64)


def single_path(N, seed, D=2, R=2):
    """Configs 1, 2, 4, 5: one path -> X (N, D), Y (N, R), theta (D+2,)."""
    rng = np.random.default_rng(seed)
    xs, ys, ts = make_path(N, rng)
    cols = [xs, ys] + ([ts] if D == 3 else [])
    X = np.ascontiguousarray(np.stack(cols, axis=1))
    return X, path_targets(xs, ys, ts, R), default_theta(D)


def batched_paths(B, N, seed=3, D=3, R=2, first=0):
    """Config 3: paths ``first .. first+B-1`` of the batch; path b uses ``default_rng([seed, b])``.

    Returns Xb (B, N, D), Yb (B, N, R), theta (D+2,).  ``first`` lets each rank generate only its shard.
    """
    Xb = np.empty((B, N, D)); Yb = np.empty((B, N, R))
    for i in range(B):
        rng = np.random.default_rng([seed, first + i])
        xs, ys, ts = make_path(N, rng)
        cols = [xs, ys] + ([ts] if D == 3 else [])
        Xb[i] = np.stack(cols, axis=1)
        Yb[i] = path_targets(xs, ys, ts, R)
    return Xb, Yb, default_theta(D)


def sweep_thetas(D=2, n_ls=8, n_noise=8):
    """Config 5 sweep: 8 lengthscales log-spaced 2e3..3e4 x 8 noise variances log-spaced 1e-3..1e-1."""
    ls = np.geomspace(2.0e3, 3.0e4, n_ls)
    nv = np.geomspace(1.0e-3, 1.0e-1, n_noise)
    out = np.empty((n_ls * n_noise, D + 2))
    k = 0
    for l in ls:
        for v in nv:
            out[k, :2] = l
            if D == 3:
                out[k, 2] = LENGTHSCALE_T
            out[k, D] = SIGNAL_VAR
            out[k, D + 1] = v
            k += 1
    return out


CONFIGS = {
    # name: (N, D, G, seed)
    "cfg1": dict(N=200, D=2, G=100, seed=1),
    "cfg2": dict(N=4096, D=2, G=512, seed=2),
    "cfg3": dict(B=4096, N=512, D=3, seed=3),
    "cfg4": dict(N=16384, D=2, seed=4),
    "cfg5": dict(N=16384, D=2, G=2048, seed=5),
}


def trajectory_families(P, k, n=33, seed=5, jitter=800.0):
    """P synthetic fixed-length 2-D trajectories in k families (a random straight segment per family in the reference's
    +-5e4 box, per-sample Gaussian jitter): xs, ys, ts of shape (P, n) -- the input of trajectories.kmeansclustering
    (GPmap.py:36-121), which resamples every trajectory to 33 points (GPmap.py:189)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-4.0e4, 4.0e4, (k, 2)); b = rng.uniform(-4.0e4, 4.0e4, (k, 2))
    fam = rng.integers(0, k, P)
    t = np.linspace(0.0, 1.0, n)
    xs = a[fam, 0:1] + (b[fam, 0:1] - a[fam, 0:1]) * t + rng.normal(0.0, jitter, (P, n))
    ys = a[fam, 1:2] + (b[fam, 1:2] - a[fam, 1:2]) * t + rng.normal(0.0, jitter, (P, n))
    ts = np.tile(np.arange(n, dtype=np.float64), (P, 1))
    return xs, ys, ts
