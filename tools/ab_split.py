"""Experiment: the batched tiled pipeline as S sub-batches on S CUDA streams (kernels of different sub-batches, bound by
different units -- covariance: FP64 issue / HBM, potf2: shared-memory wavefronts, half-tile GEMM: DMMA, backward solve:
HBM -- may overlap) against one whole-batch call:  python tools/ab_split.py [B] [N] [S ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
splits = [int(s) for s in sys.argv[3:]] or [1, 2, 3, 4]
Xb, Yb, th = wl.batched_paths(B, N, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
R = Yd.shape[2] if Yd.ndim == 3 else 1
alpha = torch.empty((B, N, R), dtype=torch.float64, device="cuda")
lml = torch.empty((B, R), dtype=torch.float64, device="cuda")
main = torch.cuda.current_stream()
ref = None
for S in splits:
    streams = [torch.cuda.Stream() for _ in range(S)]
    cuts = [B * i // S for i in range(S + 1)]

    def call():
        ev = torch.cuda.Event()
        ev.record(main)
        for i, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                lo, hi = cuts[i], cuts[i + 1]
                GPmap.fit_gp_batched(Xd[lo:hi], Yd[lo:hi], theta=th, check=False, out=(alpha[lo:hi], lml[lo:hi]))
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)

    with _native.option("path_fused", 0):
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None:
        ref = (alpha.clone(), lml.clone())
    same = torch.equal(alpha, ref[0]) and torch.equal(lml, ref[1])
    print(f"S={S}: B={B} N={N}: {ms:.3f} ms = {B / ms * 1e3:.0f} fits/s  bitwise same as S={splits[0]}: {same}", flush=True)
