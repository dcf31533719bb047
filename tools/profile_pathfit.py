"""One gpm_fit_batched call (4096 paths x N=512) forced through the one-CTA-per-path kernel (option path_fused = 2),
for `ncu --set full -k regex:path_fit`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
Xb, Yb, th = wl.batched_paths(4096, 512, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
with _native.option("path_fused", 2):
    for _ in range(3):
        a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(l).all()))
