"""One pass over every non-GEMM kernel of the path plus the N=16384 factorisation, for `ncu --set full` captures:
full covariance + fit at N (cov_kernel, potf2_inv_kernel, gemm_nt_kernel, solve kernels), then on the cfg2 model the
mean-only prediction (predict_mean_kernel), the fused cross-covariance + mean (cross_cov_mean_kernel) and the plain
cross-covariance (cross_cov_t_kernel) on a 128x128 grid."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=16384); ap.add_argument("--g", type=int, default=128)
a = ap.parse_args()
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
# large model: full covariance (both triangles: the 8 N^2-byte roofline case), then the fit
X, Y, th = wl.single_path(a.n, 4, 2, 2)
Xd = torch.from_numpy(X).cuda(); ld = (a.n + 15) // 16 * 16
K = torch.empty((a.n, ld), dtype=torch.float64, device="cuda")
_native.check(lib.gpm_cov(h, p(Xd), a.n, 2, _native.theta_array(th), p(K), ld, 0, st), "cov")
torch.cuda.synchronize(); del K
m = GPmap.fit_gp(X, Y, theta=th)
print("fit N=%d: info=%d lml=%s" % (a.n, int(m.info.item()), m.lml))
del m
# cfg2 model: the three prediction kernels
X2, Y2, th2 = wl.single_path(4096, 2, 2, 2)
m2 = GPmap.fit_gp(X2, Y2, theta=th2)
mu = m2.predict_grid(wl.BOX, (a.g, a.g), return_var=False)
mu2, var = m2.predict_grid(wl.BOX, (a.g, a.g))
Xs = torch.rand((a.g * a.g, 2), dtype=torch.float64, device="cuda") * 1e5 - 5e4
KsT = torch.empty((a.g * a.g, 4096), dtype=torch.float64, device="cuda")
_native.check(lib.gpm_cross_cov(h, p(m2.X), 4096, 2, _native.theta_array(th2), p(Xs), None, 0, a.g * a.g, p(KsT), 4096, st), "cross_cov")
torch.cuda.synchronize()
print("predict ok", float(mu.abs().max()), float(var.min()))
