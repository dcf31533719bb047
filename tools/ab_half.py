"""A/B of the whole-batch tiled pipeline with the half-tile GEMM (two CTAs per SM) and with the 128 x 128 tile kernel:
python tools/ab_half.py [B] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Xb, Yb, th = wl.batched_paths(B, N, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
res = {}
with _native.option("path_fused", 0):
    for name, val in (("half tiles", 0), ("128x128 tiles", 1)):
        with _native.option("no_half_tiles", val):
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
            print(f"{name:14s} B={B} N={N}: {ms:.3f} ms = {B / ms * 1e3:.0f} fits/s")
a0, l0 = res["128x128 tiles"]; a1, l1 = res["half tiles"]
print(f"alpha rel diff {(a1 - a0).abs().max().item() / a0.abs().max().item():.2e}, lml rel diff {((l1 - l0).abs() / l0.abs()).max().item():.2e}")
