"""Phase breakdown of path_fit_kernel (thread 0 of CTA 0, cycles summed over its paths); needs a
-DGPM_PATHFIT_TIMING build:  GPM_EXTRA_NVCC_FLAGS=-DGPM_PATHFIT_TIMING python -m gaussianprocesspathmodelling_b200.build --force"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
lib = _native.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Xb, Yb, th = wl.batched_paths(B, 512, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False); torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib.gpm_debug_pathfit_cycles(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False); e1.record(); torch.cuda.synchronize()
lib.gpm_debug_pathfit_cycles(buf, 0)
m = np.frombuffer(buf, dtype=np.int64).astype(np.float64)
npaths = len(range(0, B, min(B, 296)))
names = ["prologue", "diag update loop", "diag epilogue (cov -> packed)", "potf2 factor", "potf2 invert", "fwd z", "inv store + seg begin",
         "off U loop", "R epilogue (cov -> R buf)", "D loop", "D epilogue (L store + fwd)", "column end", "backward + out"]
print(f"B={B}: {e0.elapsed_time(e1):.3f} ms; CTA 0 handled {npaths} paths; cycles per path:")
for n, v in zip(names, m[:13]):
    print(f"  {n:32s} {v / npaths:10.0f}  ({v / npaths / 1965:.1f} us)")
print(f"  {'total':32s} {m[:13].sum() / npaths:10.0f}  ({m[:13].sum() / npaths / 1965:.1f} us)")
