"""A/B of the two batched-fit pipelines at 4096 x N=512 (one CTA per path vs whole-batch tiled launches):
python tools/ab_batched.py [B] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Xb, Yb, th = wl.batched_paths(B, N, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
res = {}
for name, val in (("path_fused", 2), ("tiled", 0)):
    with _native.option("path_fused", val):
        for _ in range(2):
            a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name] = (a.clone(), l.clone())
        print(f"{name:10s} B={B} N={N}: {ms:.3f} ms = {B / ms * 1e3:.0f} fits/s")
da = (res["path_fused"][0] - res["tiled"][0]).abs().max().item() / res["tiled"][0].abs().max().item()
dl = ((res["path_fused"][1] - res["tiled"][1]).abs() / res["tiled"][1].abs()).max().item()
print(f"alpha rel diff {da:.2e}, lml rel diff {dl:.2e}")
