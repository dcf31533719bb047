"""Per-tile phase stamps of the fused-forward panel GEMM (needs a -DGPM_GEMM_TIMING build); prints the LAST such launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Xb, Yb, th = wl.batched_paths(B, 256, seed=3)      # nblk = 2: exactly one panel launch with one tile per path
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
for _ in range(2): GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
lib.gpm_debug_gemm_marks(buf)
m = np.frombuffer(buf, dtype=np.int64).reshape(2, 4, 16).astype(np.float64)
for w, name in ((0, "warp 0"), (1, "warp 7")):
    t = m[w, 0]
    print(name, "main loop %.0f | store epilogue %.0f | partial sums %.0f | barrier wait %.0f | reduce+atomics %.0f" % (
        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]))
