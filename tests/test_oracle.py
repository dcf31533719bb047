"""Pin the numpy/scipy oracle (parity is unpinned by the reference, which has no GP code):
hand-computed posteriors, closed-form limits, scikit-learn, mpmath at 50 digits, properties."""
import math

import numpy as np
import pytest

from oracle import gp_ref
from gaussianprocesspathmodelling_b200 import workloads as wl


def test_n1_posterior_by_hand():
    X = np.array([[3.0, -1.0]]); Y = np.array([[2.0]]); th = np.array([2.0, 2.0, 1.5, 0.25])
    m = gp_ref.fit(X, Y, th)
    k = 1.5 + 0.25
    assert np.allclose(m["L"], math.sqrt(k))
    assert np.allclose(m["alpha"], 2.0 / k)
    assert np.allclose(m["lml"], -0.5 * 4.0 / k - 0.5 * math.log(k) - 0.5 * math.log(2 * math.pi))
    xs = np.array([[4.0, 1.0]])
    ks = 1.5 * math.exp(-0.5 * ((1.0 / 2.0) ** 2 + (2.0 / 2.0) ** 2))
    mu, var = gp_ref.predict(m, xs)
    assert np.allclose(mu, ks * 2.0 / k)
    assert np.allclose(var, 1.5 - ks * ks / k)


def test_n2_posterior_by_hand():
    X = np.array([[0.0, 0.0], [1.0, 0.0]]); Y = np.array([[1.0], [-1.0]]); th = np.array([1.0, 1.0, 1.0, 0.1])
    r = math.exp(-0.5)
    K = np.array([[1.1, r], [r, 1.1]])
    Kinv = np.array([[1.1, -r], [-r, 1.1]]) / (1.1 ** 2 - r ** 2)
    m = gp_ref.fit(X, Y, th)
    assert np.allclose(gp_ref.cov(X, th), K)
    assert np.allclose(m["alpha"], Kinv @ Y)
    lml = -0.5 * float(Y.T @ Kinv @ Y) - 0.5 * math.log(1.1 ** 2 - r ** 2) - math.log(2 * math.pi)
    assert np.allclose(m["lml"], lml)
    xs = np.array([[0.5, 0.0]])
    ks = np.array([math.exp(-0.125), math.exp(-0.125)])
    mu, var = gp_ref.predict(m, xs)
    assert np.allclose(mu[0, 0], ks @ Kinv @ Y[:, 0])
    assert np.allclose(var[0], 1.0 - ks @ Kinv @ ks)


def test_iid_limit_of_lml():
    # lengthscale -> 0: K -> (sf2 + sn2) I, LML = sum log N(y; 0, sf2 + sn2)
    rng = np.random.default_rng(0)
    X = rng.uniform(-5e4, 5e4, (40, 2)); Y = rng.standard_normal((40, 2))
    th = np.array([1e-3, 1e-3, 0.7, 0.3])
    lml = gp_ref.fit(X, Y, th)["lml"]
    expect = -0.5 * (Y ** 2).sum(0) / 1.0 - 0.5 * 40 * math.log(2 * math.pi * 1.0)
    assert np.allclose(lml, expect, rtol=1e-12)


def test_cov_symmetric_psd_and_translation_invariant():
    X, _, th = wl.single_path(200, seed=1)
    K = gp_ref.cov(X, th)
    assert np.array_equal(K, K.T)
    assert np.linalg.eigvalsh(K).min() > 0
    K2 = gp_ref.cov(X + np.array([123.0, -77.0]), th)
    assert np.allclose(K, K2, rtol=1e-12)


def test_lml_permutation_invariant_and_variance_bounds():
    X, Y, th = wl.single_path(150, seed=5)
    p = np.random.default_rng(1).permutation(150)
    a = gp_ref.fit(X, Y, th); b = gp_ref.fit(X[p], Y[p], th)
    assert np.allclose(a["lml"], b["lml"], rtol=1e-11)
    mu, var = gp_ref.predict_grid(a, wl.BOX, (20, 20))
    assert var.min() > -1e-9 and var.max() <= th[2] + 1e-12


def test_against_sklearn(golden):
    # independent implementation: LML, mean, variance on config 1 (N=200, 100x100 grid)
    assert abs(golden["cfg1_lml"][0] - golden["cfg1_sk_lml0"]) < 1e-9 * abs(golden["cfg1_sk_lml0"])
    mu, mu_sk = golden["cfg1_mu"][:, :, 0], golden["cfg1_sk_mu0"]
    assert np.abs(mu - mu_sk).max() / np.abs(mu_sk).max() < 1e-11
    var, var_sk = golden["cfg1_var"], golden["cfg1_sk_var"]
    assert np.abs(var - var_sk).max() / np.abs(var_sk).max() < 1e-10


def test_against_mpmath_50_digits():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    N = 24
    X, Y, th = wl.single_path(N, seed=21, D=2, R=1)
    ls, sf2, sn2 = th[:2], th[2], th[3]
    Xm = [[mp.mpf(float(v)) / mp.mpf(float(l)) for v, l in zip(row, ls)] for row in X]
    def k(a, b):
        return mp.mpf(float(sf2)) * mp.exp(-sum((p - q) ** 2 for p, q in zip(a, b)) / 2)
    K = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            K[i, j] = k(Xm[i], Xm[j]) + (mp.mpf(float(sn2)) if i == j else 0)
    y = mp.matrix([mp.mpf(float(v)) for v in Y[:, 0]])
    alpha = mp.lu_solve(K, y)
    L = mp.cholesky(K)
    lml = -(y.T * alpha)[0] / 2 - sum(mp.log(L[i, i]) for i in range(N)) - mp.mpf(N) / 2 * mp.log(2 * mp.pi)
    m = gp_ref.fit(X, Y, th)
    cond = np.linalg.cond(gp_ref.cov(X, th))
    a_ref = np.array([float(v) for v in alpha])
    assert np.abs(m["alpha"][:, 0] - a_ref).max() / np.abs(a_ref).max() < 50 * cond * 2.2e-16
    assert abs(m["lml"][0] - float(lml)) < 1e-11 * abs(float(lml))
    xs = [mp.mpf(1234.0) / mp.mpf(float(ls[0])), mp.mpf(-4321.0) / mp.mpf(float(ls[1]))]
    ks = mp.matrix([k(xs, Xm[i]) for i in range(N)])
    mu = (ks.T * alpha)[0]
    var = mp.mpf(float(sf2)) - (ks.T * mp.lu_solve(K, ks))[0]
    mu_o, var_o = gp_ref.predict(m, np.array([[1234.0, -4321.0]]))
    assert abs(mu_o[0, 0] - float(mu)) < 1e-9 * max(1.0, abs(float(mu)))
    assert abs(var_o[0] - float(var)) < 1e-9


def test_grid_points_layout():
    P = gp_ref.grid_points((0.0, 3.0, 10.0, 12.0), (4, 3))
    assert P.shape == (12, 2)
    assert np.array_equal(P[:4, 0], [0.0, 1.0, 2.0, 3.0]) and np.all(P[:4, 1] == 10.0)
    assert P[4, 1] == 11.0 and P[-1].tolist() == [3.0, 12.0]
    P3 = gp_ref.grid_points((0.0, 1.0, 0.0, 1.0), (2, 2), t=5.0)
    assert P3.shape == (4, 3) and np.all(P3[:, 2] == 5.0)


def test_batched_and_sweep_wrappers():
    Xb, Yb, th = wl.batched_paths(3, 40, seed=3)
    a, l = gp_ref.fit_batched(Xb, Yb, th)
    m1 = gp_ref.fit(Xb[1], Yb[1], th)
    assert np.array_equal(a[1], m1["alpha"]) and np.array_equal(l[1], m1["lml"])
    ths = wl.sweep_thetas(D=3)[:3]
    s = gp_ref.lml_sweep(Xb[0], Yb[0], ths)
    assert s.shape == (3, 2) and np.allclose(s[2], gp_ref.fit(Xb[0], Yb[0], ths[2])["lml"])


def test_reference_restatements_match_the_reference(ref_golden):
    # the code the reference DOES contain, pinned by vectors generated from the reference itself
    g = ref_golden
    xs, ys, ts = g["xs"], g["ys"], g["ts"]
    P = xs.shape[0]
    for a in range(P):
        for b in range(P):
            assert gp_ref.calc_distance(xs[a], ys[a], xs[b], ys[b]) == g["dist"][a, b]
    for p in range(P):
        assert gp_ref.travel_sum(xs[p], ys[p]) == g["sums"][p]
    sizes = g["group_sizes"]; o = 0
    for c, n in enumerate(sizes):
        mx, my, mt = gp_ref.calc_mean_traj(xs[o:o + n], ys[o:o + n], ts[o:o + n])
        assert np.array_equal(mx, g["cx"][c]) and np.array_equal(my, g["cy"][c]) and np.array_equal(mt, g["ct"][c])
        o += n
    for p in range(P):
        for c in range(len(sizes)):
            # reference order: calc_distance(centroid, path).  Fractional centroids: the reference's
            # np.linalg.norm uses BLAS ddot (FMA on this CPU), so agreement is to an ulp or two.
            d = gp_ref.calc_distance(g["cx"][c], g["cy"][c], xs[p], ys[p])
            assert abs(d - g["d2c"][p, c]) <= 1e-15 * g["d2c"][p, c]


def test_lml_grad_matches_finite_differences():
    X, Y, th = wl.single_path(60, seed=8, D=3, R=2)
    g = gp_ref.lml_grad(X, Y, th)
    assert g.shape == (2, 5)
    for j in range(5):
        h = 1e-5
        tp, tm = th.copy(), th.copy()
        tp[j] *= np.exp(h); tm[j] *= np.exp(-h)
        fd = (gp_ref.fit(X, Y, tp)["lml"] - gp_ref.fit(X, Y, tm)["lml"]) / (2 * h)
        assert np.allclose(g[:, j], fd, rtol=2e-6, atol=1e-6 * np.abs(g).max())


def test_separable_form_of_the_rbf_kernel_agrees_to_argument_rounding():
    """The grid kernels evaluate k = exp(-dx^2/2) * (sf2 exp(-dy^2/2)) instead of sf2 exp(-(dx^2+dy^2)/2)
    (csrc/cov.cu: cross_cov_grid_kernel, predict_mean_grid_kernel).  In exact arithmetic they are equal; in float64
    each form carries the rounding of its exponent(s), |arg| * eps relative, so the two agree to (|arg| + 4) eps
    relative (1.5e-14 at |arg| = 150, where k ~ 1e-65) and to a few eps * sf2 in absolute terms -- against tolerances of
    1e-8 (mean) and 1e-6 (variance)."""
    rng = np.random.default_rng(0)
    dx = rng.uniform(-12.5, 12.5, 200000)          # scaled distances of the workloads: |d| <= 100/8 per axis
    dy = rng.uniform(-12.5, 12.5, 200000)
    sf2 = 1.7
    arg = 0.5 * (dx * dx + dy * dy)
    joint = sf2 * np.exp(-arg)
    sep = np.exp(-0.5 * (dx * dx)) * (sf2 * np.exp(-0.5 * (dy * dy)))
    eps = np.finfo(np.float64).eps
    assert np.all(np.abs(sep - joint) <= (arg + 4.0) * eps * joint)
    assert np.abs(sep - joint).max() <= 4 * eps * sf2
