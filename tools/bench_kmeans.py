"""The reference's real hot loop -- trajectories.kmeansclustering (GPmap.py:36-121) -- on the device: time per Lloyd
iteration (assignment + centroid update + convergence sum) for P synthetic 33-point trajectories, k clusters, against
the numpy restatement of the reference's loops (oracle/gp_ref.py: lloyd) on a bounded sample of the paths.
usage: python tools/bench_kmeans.py [P] [k]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import _native, workloads as wl
from oracle import gp_ref
P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = 33
rng = np.random.default_rng(5)
xs, ys, ts = wl.trajectory_families(P, k, n, seed=5)
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda x: C.c_void_p(x.data_ptr())
dev = torch.from_numpy(np.stack([xs, ys, ts])).cuda()
pxT, pyT = dev[0].t().contiguous(), dev[1].t().contiguous()
init = rng.choice(P, k, replace=False)
def run(iters):
    cents = dev[:, torch.from_numpy(init).cuda(), :].contiguous()
    assign = torch.zeros(P, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.gpm_kmeans_workspace_bytes(P, n, k)) // 8, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # threshold 0: never converges early, every iteration does its full work
    _native.check(lib.gpm_kmeans_lloyd(h, p(dev[0]), p(dev[1]), p(dev[2]), p(pxT), p(pyT), P, n, k, p(cents), p(assign), 0.0, iters, 1, p(ws), st), "lloyd")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, assign
run(2)
ms, assign = run(10)
print(f"GPU: P={P} n={n} k={k}: {ms:.3f} ms per Lloyd iteration = {P / ms * 1e3 / 1e6:.1f} M path assignments/s; "
      f"largest cluster {int(torch.bincount(assign.long(), minlength=k).max())} paths (its ordered sum is a serial chain)")
Ps = min(P, 1500)
t0 = time.perf_counter()
a_o, c_o, it = gp_ref.lloyd(xs[:Ps], ys[:Ps], ts[:Ps], list(range(k)), threshold=0.0, max_iter=1)
dt = time.perf_counter() - t0
print(f"CPU (numpy restatement of the reference's loops, 1 iteration on {Ps} paths): {dt * 1e3:.1f} ms = {Ps / dt / 1e3:.1f} k path assignments/s")
