"""One batched call (4096 paths x N=512) for ncu / timing: python tools/profile_batched.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Xb, Yb, th = wl.batched_paths(4096, 512, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
for _ in range(2):
    GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"batched 4096 x 512: {ms:.3f} ms = {4096 / ms * 1e3:.0f} fits/s, finite={bool(torch.isfinite(l).all())}")
