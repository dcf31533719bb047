"""Property tests (hypothesis): invariants of the GP path that hold for any input, checked on the
oracle (CPU) and on the CUDA path (GPU)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import gp_ref
from gaussianprocesspathmodelling_b200 import workloads as wl


def make_case(n, d, seed):
    rng = np.random.default_rng(seed)
    X = np.rint(rng.uniform(-5e4, 5e4, (n, d)))
    if d == 3:
        X[:, 2] = 0.5 * np.arange(n)
    Y = rng.standard_normal((n, 2))
    th = np.array([rng.uniform(3e3, 2e4)] * 2 + ([rng.uniform(1.0, 20.0)] if d == 3 else []) + [rng.uniform(0.5, 2.0), rng.uniform(1e-3, 1e-1)])
    return X, Y, th


@settings(max_examples=15, deadline=None)
@given(n=st.integers(2, 60), d=st.sampled_from([2, 3]), seed=st.integers(0, 10 ** 6))
def test_oracle_invariants(n, d, seed):
    X, Y, th = make_case(n, d, seed)
    m = gp_ref.fit(X, Y, th)
    p = np.random.default_rng(seed + 1).permutation(n)
    mp = gp_ref.fit(X[p], Y[p], th)
    assert np.allclose(m["lml"], mp["lml"], rtol=1e-10, atol=1e-9)               # permutation invariance
    shift = np.zeros(d); shift[:2] = [1234.0, -987.0]
    ms = gp_ref.fit(X + shift, Y, th)
    assert np.allclose(m["lml"], ms["lml"], rtol=1e-10, atol=1e-9)               # translation invariance
    mu, var = gp_ref.predict(m, X)                                               # at the training inputs
    assert var.min() > -1e-9 and var.max() <= th[d] + 1e-12
    K = gp_ref.cov(X, th)
    assert np.allclose(K @ m["alpha"], Y, atol=1e-8 * max(1.0, np.abs(Y).max()) * np.linalg.cond(K) ** 0.5)
    # linearity of alpha and the mean in the targets
    m2 = gp_ref.fit(X, 3.0 * Y, th)
    assert np.allclose(m2["alpha"], 3.0 * m["alpha"], rtol=1e-12)


@pytest.mark.gpu
@settings(max_examples=12, deadline=None)
@given(n=st.integers(1, 400), d=st.sampled_from([2, 3]), m=st.integers(1, 300), seed=st.integers(0, 10 ** 6))
def test_gpu_matches_oracle_on_random_cases(n, d, m, seed):
    from gaussianprocesspathmodelling_b200 import GPmap
    X, Y, th = make_case(n, d, seed)
    rng = np.random.default_rng(seed + 7)
    Xs = np.column_stack([rng.uniform(-5e4, 5e4, m), rng.uniform(-5e4, 5e4, m)] + ([rng.uniform(0, 0.5 * n, m)] if d == 3 else []))
    mo = gp_ref.fit(X, Y, th)
    mu_o, var_o = gp_ref.predict(mo, Xs)
    g = GPmap.fit_gp(X, Y, theta=th)
    mu, var = g.predict(Xs)
    scale = max(np.abs(mo["alpha"]).max(), 1e-300)
    assert np.abs(g.alpha.cpu().numpy() - mo["alpha"]).max() / scale < 1e-8
    assert np.abs(g.lml - mo["lml"]).max() < 1e-10 * max(1.0, np.abs(mo["lml"]).max())
    assert np.abs(mu.cpu().numpy() - mu_o).max() <= 1e-8 * max(np.abs(mu_o).max(), 1e-12)
    assert np.abs(var.cpu().numpy() - var_o).max() <= 1e-6 * max(np.abs(var_o).max(), 1e-12)
    v = var.cpu().numpy()
    assert v.min() > -1e-9 and v.max() <= th[d] + 1e-12
