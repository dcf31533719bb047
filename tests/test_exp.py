"""The hand-rolled exp(x), x <= 0, of the covariance kernels (csrc/exp_neg.cuh) on the CPU: the header compiles as
plain C with libm's fma(), every operation in it is a correctly rounded IEEE operation written out explicitly, so
the device code computes the same bits.  All three table sizes are checked against mpmath (<= 0.56 ulp; 0.52 for the default J = 16) and numpy (<= 1 ulp), at the underflow
edge, and for arguments large enough to wrap the integer exponent if it were not screened."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "gaussianprocesspathmodelling_b200", "csrc", "exp_neg.cuh")


@pytest.fixture(scope="module", params=[4, 3, 6], ids=["J16-default", "J8", "J64"])
def lib(tmp_path_factory, request):
    d = tmp_path_factory.mktemp("expneg")
    src = d / "shim.c"
    src.write_text(f'#include "{HDR}"\n'
                   "void v_exp_neg(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = gpm_exp_neg(x[i]); }\n"
                   "void v_exp_neg_half(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = gpm_exp_neg_half(x[i]); }\n")
    so = d / "shim.so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", f"-DGPM_EXP_LOG2J={request.param}", "-shared", "-fPIC", "-x", "c",
                    "-o", str(so), str(src), "-lm"], check=True)
    L = ctypes.CDLL(str(so))
    for f in (L.v_exp_neg, L.v_exp_neg_half):
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return L


def _run(f, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    f(x.ctypes.data, y.ctypes.data, len(x))
    return y


def test_exp_neg_within_one_ulp_of_numpy_and_half_an_ulp_of_mpmath(lib):
    import mpmath as mp
    mp.mp.prec = 120
    rng = np.random.default_rng(0)
    x = -np.concatenate([rng.uniform(0, 50, 300000), rng.uniform(0, 707, 200000), np.exp(rng.uniform(-40, 3, 100000))])
    for f, arg, ref in ((lib.v_exp_neg, x, np.exp(x)), (lib.v_exp_neg_half, -2.0 * x, np.exp(x))):
        y = _run(f, arg)
        assert np.max(np.abs(y - ref) / np.spacing(ref)) <= 1.0
        worst = 0.0
        for i in rng.choice(len(x), 4000, replace=False):
            t = mp.exp(mp.mpf(float(x[i])))
            worst = max(worst, float(abs(mp.mpf(float(y[i])) - t) / mp.mpf(float(np.spacing(float(t))))))
        assert worst <= 0.56, worst          # J=16 (default): 0.52; J=8 carries a larger e^t - 1 term: 0.55


def test_exp_neg_edges(lib):
    x = np.array([0.0, -0.0, -1e-300, 1e-17, -707.0, -707.7, -708.5, -745.0, -1e4, -5e7, -1e300, -np.inf])
    y = _run(lib.v_exp_neg, x)
    assert y[0] == 1.0 and y[1] == 1.0 and y[2] == 1.0 and y[3] == 1.0
    assert y[4] == np.exp(-707.0)
    assert np.all(y[6:] == 0.0)                                  # flushed, never garbage from a wrapped exponent
    assert np.all(np.abs(y - np.exp(x)) < 4.5e-308)              # the flush costs less than the smallest normal number
    d2 = np.array([0.0, 1e-300, 2.0, 1414.0, 1417.0, 1e9, 1e300, np.inf])
    z = _run(lib.v_exp_neg_half, d2)
    assert z[0] == 1.0 and z[1] == 1.0 and abs(z[2] - np.exp(-1.0)) <= np.spacing(np.exp(-1.0)) and z[3] == np.exp(-707.0)
    assert np.all(z[4:] == 0.0)
    # monotone across the table-index boundaries (k ln2/64)
    xs = -np.linspace(0.0, 3.0, 200001)
    ys = _run(lib.v_exp_neg, xs)
    assert np.all(np.diff(ys) <= 0.0)
