"""Per-tile phase stamps (cycles) of the factorisation's tile GEMM launches in one batched fit (CTA x=0 of path 7);
needs a -DGPM_GEMM_TIMING build (GPM_EXTRA_NVCC_FLAGS=-DGPM_GEMM_TIMING python -m gaussianprocesspathmodelling_b200.build --force)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Xb, Yb, th = wl.batched_paths(B, N, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
buf = (C.c_longlong * 512)()
GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
lib.gpm_debug_gemm_marks(buf)          # also resets the launch counter
GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
lib.gpm_debug_gemm_marks(buf)
m = np.frombuffer(buf, dtype=np.int64).reshape(8, 2, 4, 8).astype(np.float64)
for L in range(8):
    if m[L, 0, 0, 0] == 0:
        continue
    for w, name in ((0, "warp0"), (1, "warp7")):
        t = m[L, w]
        out = [f"launch {L} {name}: entry->tile0 {t[0, 0] - t[0, 7]:.0f}"]
        for k in range(4):
            if t[k, 0] == 0 or (k > 0 and t[k, 0] < t[k - 1, 0]):
                break
            cw = f" (C wait {t[k, 6] - t[k, 1]:.0f})" if t[k, 6] > t[k, 1] else ""
            fw = f" fwd {t[k, 5] - t[k, 2]:.0f}" if t[k, 5] > t[k, 2] else ""
            gap = f" gap {t[k, 0] - max(t[k - 1, 2], t[k - 1, 5]):.0f}" if k > 0 else ""
            out.append(f"|{gap} main {t[k, 1] - t[k, 0]:.0f}{cw} epi {t[k, 2] - t[k, 1]:.0f}{fw}")
        print(" ".join(out))
