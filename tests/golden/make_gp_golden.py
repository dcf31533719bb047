"""Golden GP vectors.  PARITY IS UNPINNED BY THE REFERENCE (it has no GP code):
these vectors come from this repo's numpy/scipy oracle (oracle/gp_ref.py) and, independently, from
scikit-learn's GaussianProcessRegressor; both are stored so the tests can check either against the
CUDA path and against each other.  Outputs: tests/golden/gp_golden.npz (committed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gp_ref                                              # noqa: E402
from gaussianprocesspathmodelling_b200 import workloads as wl          # noqa: E402


def sklearn_fit_predict(X, y, theta, Xs):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as Cst
    D = X.shape[1]
    ls, sf2, sn2 = theta[:D], theta[D], theta[D + 1]
    gpr = GaussianProcessRegressor(kernel=Cst(sf2, "fixed") * RBF(ls, "fixed"), alpha=sn2, optimizer=None)
    gpr.fit(X, y)
    mu, sd = gpr.predict(Xs, return_std=True)
    return gpr.log_marginal_likelihood_value_, mu, sd ** 2


def main():
    out = {}
    # (a) the reference's own path length: N=33, D=2 and D=3, small grid
    for tag, D in (("n33d2", 2), ("n33d3", 3)):
        X, Y, th = wl.single_path(33, seed=11, D=D, R=2)
        m = gp_ref.fit(X, Y, th)
        t = 7.5 if D == 3 else None
        mu, var = gp_ref.predict_grid(m, wl.BOX, (9, 7), t=t)
        out.update({f"{tag}_X": X, f"{tag}_Y": Y, f"{tag}_theta": th, f"{tag}_K": gp_ref.cov(X, th), f"{tag}_L": m["L"],
                    f"{tag}_alpha": m["alpha"], f"{tag}_lml": m["lml"], f"{tag}_mu": mu, f"{tag}_var": var})
    # (b) config 1: N=200, 100x100 grid (the reference's CPU-runnable case)
    X, Y, th = wl.single_path(200, seed=1, D=2, R=2)
    m = gp_ref.fit(X, Y, th)
    mu, var = gp_ref.predict_grid(m, wl.BOX, (100, 100))
    lml_sk, mu_sk, var_sk = sklearn_fit_predict(X, Y[:, 0], th, gp_ref.grid_points(wl.BOX, (100, 100)))
    out.update({"cfg1_alpha": m["alpha"], "cfg1_lml": m["lml"], "cfg1_mu": mu, "cfg1_var": var,
                "cfg1_sk_lml0": np.array(lml_sk), "cfg1_sk_mu0": mu_sk.reshape(100, 100), "cfg1_sk_var": var_sk.reshape(100, 100)})
    # (c) ragged size crossing two 128-blocks: N=300, D=3, scattered query points
    X, Y, th = wl.single_path(300, seed=12, D=3, R=1)
    rng = np.random.default_rng(13)
    Xs = np.column_stack([rng.uniform(-5e4, 5e4, 257), rng.uniform(-5e4, 5e4, 257), rng.uniform(0, 150, 257)])
    m = gp_ref.fit(X, Y, th)
    mu, var = gp_ref.predict(m, Xs)
    out.update({"n300_Xs": Xs, "n300_alpha": m["alpha"], "n300_lml": m["lml"], "n300_mu": mu, "n300_var": var})
    # (d) batched: 6 paths x N=130, D=3
    Xb, Yb, th = wl.batched_paths(6, 130, seed=3, D=3, R=2)
    a, l = gp_ref.fit_batched(Xb, Yb, th)
    out.update({"b6_alpha": a, "b6_lml": l})
    np.savez_compressed(os.path.join(HERE, "gp_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print("cfg1 lml oracle", m["lml"], "sk", lml_sk)


if __name__ == "__main__":
    main()
