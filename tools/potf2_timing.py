"""Phase breakdown of potf2_inv_kernel (CTA 0) from clock64 stamps; needs a -DGPM_POTF2_TIMING build.
usage: potf2_timing.py [N] (N=128: lone block, 16 warps)  or  potf2_timing.py batch (4096 x 512: 8 warps, 3 CTAs/SM)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load()
if len(sys.argv) > 1 and sys.argv[1] == "batch":
    Xb, Yb, th = wl.batched_paths(4096, 512, seed=3)
    Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
    for _ in range(2): GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
else:
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    X, Y, th = wl.single_path(N, 4, 2, 2)
    for _ in range(2): GPmap.fit_gp(X, Y, theta=th)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
lib.gpm_debug_potf2_marks(buf)
m = np.frombuffer(buf, dtype=np.int64).astype(np.float64)
print("load %.0f, first barrier %.0f, first 8x8 factor %.0f" % (m[0] - m[58], m[1] - m[0], m[2] - m[1]))
fw = np.array([m[3 + 3 * p] - m[2 + 3 * p] for p in range(16)])
up = np.array([m[4 + 3 * p] - m[3 + 3 * p] for p in range(15)])
print("fwdsub  per panel:", fw.astype(int).tolist(), "sum", int(fw.sum()))
print("update+lookahead factor per panel:", up.astype(int).tolist(), "sum", int(up.sum()))
print("look-ahead lane: barrier -> start of the 8x8 factorisation:", [int(m[64 + 2 * p] - m[3 + 3 * p]) for p in range(15)])
print("look-ahead lane: 8x8 factorisation:", [int(m[65 + 2 * p] - m[64 + 2 * p]) for p in range(15)])
print("write L %.0f | inv level0 %.0f | levels 8/16/32/64: %.0f %.0f %.0f %.0f | fused z %.0f | store inv %.0f" % (
    m[50] - m[48], m[51] - m[50], m[52] - m[51], m[53] - m[52], m[54] - m[53], m[55] - m[54], m[57] - m[55], m[56] - m[57]))
print("total from end of load to end: %.0f cycles" % (m[56] - m[0]))
