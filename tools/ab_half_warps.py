"""A/B of the half-tile GEMM with 8 warps per CTA (32 x 32 warp tiles, four warps per scheduler) and with 4 (64 x 32):
python tools/ab_half_warps.py [B] [N] [R]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
Xb, Yb, th = wl.batched_paths(B, N, seed=3, R=min(R, 2))
if R > 2:
    Yb = np.concatenate([Yb] + [Yb[:, :, :1] * (k + 2) for k in range(R - 2)], axis=2)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
ref = None
with _native.option("path_fused", 0):
    for nw, st in ((4, 0), (8, 2), (8, 3), (8, 0)):
        with _native.option("half_warps", nw), _native.option("half_stages", st):
            for _ in range(2):
                a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                a, l = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            if ref is None: ref = (a.clone(), l.clone())
            print(f"{nw} warps per CTA, ring cap {st}  B={B} N={N} R={R}: {ms:.3f} ms = {B / ms * 1e3:.0f} fits/s, bitwise equal to 4 warps: {bool(torch.equal(a, ref[0]) and torch.equal(l, ref[1]))}", flush=True)
