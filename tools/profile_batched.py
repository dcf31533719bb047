"""Config 3 (batched fits) profile driver: warm-up + one measured call."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--b", type=int, default=4096); ap.add_argument("--n", type=int, default=512); ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
Xb, Yb, th = wl.batched_paths(a.b, a.n, seed=3)
Xd, Yd = torch.from_numpy(Xb).cuda(), torch.from_numpy(Yb).cuda()
for s in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); alpha, lml = GPmap.fit_gp_batched(Xd, Yd, theta=th, check=False); e1.record(); torch.cuda.synchronize()
    print(f"call {s}: {e0.elapsed_time(e1):.3f} ms for {a.b} fits -> {a.b / e0.elapsed_time(e1) * 1e3:.0f} fits/s")
