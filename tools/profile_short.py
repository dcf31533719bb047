"""16384 paths x N=33 (the reference's trajectory length) through fit_small_kernel, for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 33
Xb, Yb, th = wl.batched_paths(16384, N, seed=3, D=2, R=2)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
for _ in range(3):
    a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
print("ok", float(l.mean()))
