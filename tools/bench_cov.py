"""CUDA-event timings of the exp-bound kernels: covariance (full / lower tiles) at N=16384, the pointwise
cross-covariance and fused cross-cov+mean at M=16384, the mean-only kernels, the batched covariance.
usage: python tools/bench_cov.py [tag]   -> one JSON line"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: C.c_void_p(t.data_ptr())
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flushbuf.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps
out = {"tag": sys.argv[1] if len(sys.argv) > 1 else ""}
N = 16384
X, Y, th = wl.single_path(N, 4, 2, 2)
Xd = torch.from_numpy(X).cuda(); tha = _native.theta_array(th)
K = torch.empty((N, N), dtype=torch.float64, device="cuda")
for name, fl in (("cov_full", 0), ("cov_lower", 1)):
    t = timed(lambda: _native.check(lib.gpm_cov(h, ptr(Xd), N, 2, tha, ptr(K), N, fl, st), "cov"))
    out[name + "_ms"] = t
out["cov_full_tbs"] = 8.0 * N * N / out["cov_full_ms"] / 1e9
M = 16384
Xs = torch.from_numpy(np.random.default_rng(0).uniform(-5e4, 5e4, (M, 2))).cuda()
t = timed(lambda: _native.check(lib.gpm_cross_cov(h, ptr(Xd), N, 2, tha, ptr(Xs), None, 0, M, ptr(K), N, st), "cc"))
out["cross_cov_t_ms"] = t; out["cross_cov_t_tbs"] = 8.0 * N * M / t / 1e9
del K
X2, Y2, th2 = wl.single_path(4096, 2, 2, 2)
m = GPmap.fit_gp(X2, Y2, theta=th2)
Xq = torch.from_numpy(np.random.default_rng(1).uniform(-5e4, 5e4, (65536, 2))).cuda()
out["predict_pointwise_mean_ms_M65536_N4096"] = timed(lambda: m.predict(Xq, return_var=False), 3, 1)
out["predict_grid_mean_ms_512x512_N4096"] = timed(lambda: m.predict_grid(wl.BOX, (512, 512), return_var=False), 3, 1)
# fused cross-cov + mean (pointwise queries with variance): time only that kernel through predict with a tiny N? use M=16384
Xq2 = Xq[:16384].contiguous()
out["predict_pointwise_meanvar_ms_M16384_N4096"] = timed(lambda: m.predict(Xq2), 3, 1)
Xb, Yb, thb = wl.batched_paths(2048, 33, seed=3, D=2, R=2)
Xbd, Ybd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
out["short_paths_N33_B2048_ms"] = timed(lambda: GPmap.fit_gp_batched(Xbd, Ybd, theta=thb, check=False), 5, 2)
print(json.dumps(out))
