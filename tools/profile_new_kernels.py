"""cfg2 model: separable grid kernels on the full 512x512 grid (cross_cov_grid_kernel into a K*^T buffer, predict_mean_grid_kernel),
then the short-path kernel (16384 paths x N=33), for `ncu --set full` captures."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl  # noqa: E402
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
X2, Y2, th2 = wl.single_path(4096, 2, 2, 2)
m2 = GPmap.fit_gp(X2, Y2, theta=th2)
G = 512
grid = _native.GpmGrid(wl.BOX[0], wl.BOX[1], wl.BOX[2], wl.BOX[3], 0.0, G, G)
KsT = torch.empty((G * G, 4096), dtype=torch.float64, device="cuda")
for _ in range(2):
    _native.check(lib.gpm_cross_cov(h, p(m2.X), 4096, 2, _native.theta_array(th2), None, C.byref(grid), 0, G * G, p(KsT), 4096, st), "cross_cov")
    mu = m2.predict_grid(wl.BOX, (G, G), return_var=False)
torch.cuda.synchronize()
Xb, Yb, thb = wl.batched_paths(16384, 33, seed=3, D=2, R=2)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
for _ in range(2):
    a, l = GPmap.fit_gp_batched(Xd, Yd, theta=thb, check=False)
torch.cuda.synchronize()
print("ok", float(mu.abs().max()), float(l.mean()))
