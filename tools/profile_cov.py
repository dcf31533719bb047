"""The exp-bound covariance kernels once each, for `ncu --set full`: cov_kernel (lower tiles, then full) at N=16384,
cross_cov_t_kernel and cross_cov_mean_kernel (pointwise queries, M=16384) on the N=4096 model."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
N = 16384
X, Y, th = wl.single_path(N, 4, 2, 2)
Xd = torch.from_numpy(X).cuda(); tha = _native.theta_array(th)
K = torch.empty((N, N), dtype=torch.float64, device="cuda")
for fl in (1, 0):
    _native.check(lib.gpm_cov(h, p(Xd), N, 2, tha, p(K), N, fl, st), "cov")
M = 16384
Xs = torch.from_numpy(np.random.default_rng(0).uniform(-5e4, 5e4, (M, 2))).cuda()
_native.check(lib.gpm_cross_cov(h, p(Xd), N, 2, tha, p(Xs), None, 0, M, p(K), N, st), "cc")
torch.cuda.synchronize(); del K
X2, Y2, th2 = wl.single_path(4096, 2, 2, 2)
m = GPmap.fit_gp(X2, Y2, theta=th2)
mu, var = m.predict(Xs)
torch.cuda.synchronize()
print("ok", float(mu.abs().max()), float(var.min()))
