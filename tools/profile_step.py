"""One warm-up step + one measured step of config 2 (or --n/--g), for ncu launch lists and captures."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--g", type=int, default=512)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
X, Y, th = wl.single_path(a.n, 2, 2, 2)
Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
lib = _native.load()
for s in range(a.steps):
    n0 = lib.gpm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m = GPmap.fit_gp(Xd, Yd, theta=th, check=False)
    mu, var = m.predict_grid(wl.BOX, (a.g, a.g))
    e1.record(); torch.cuda.synchronize()
    print(f"step {s}: {e0.elapsed_time(e1):.3f} ms, {lib.gpm_launch_count() - n0} library launches, info={int(m.info.item())}")
