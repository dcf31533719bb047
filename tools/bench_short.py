"""Short-path fits (one CTA per path, fit_small_kernel): fits/s at the reference's path length and around it.
usage: python tools/bench_short.py [B] [N ...]   (GPM_SMALL_TWO_MAX=0: one thread per row always)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for N in ([int(v) for v in sys.argv[2:]] or (16, 32, 33, 40, 64, 80, 112)):
    Xb, Yb, th = wl.batched_paths(B, N, seed=3, D=2, R=2)
    Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
    for _ in range(3):
        a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"N={N:3d} B={B}: {ms:.4f} ms = {B / ms * 1e3 / 1e6:.1f} M fits/s")
