"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): ragged sizes, D=3, batched, kmeans."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl  # noqa: E402
X, Y, th = wl.single_path(300, 12, 3, 2)
m = GPmap.fit_gp(X, Y, theta=th)
mu, var = m.predict_grid(wl.BOX, (19, 13), t=3.0)
Xb, Yb, thb = wl.batched_paths(3, 130, seed=3)
a, l = GPmap.fit_gp_batched(Xb, Yb, theta=thb)
X2, Y2, th2 = wl.single_path(520, 5, 2, 1)
m2 = GPmap.fit_gp(X2, Y2, theta=th2)
mu2, var2 = m2.predict(np.random.default_rng(0).uniform(-5e4, 5e4, (200, 2)))
torch.cuda.synchronize()
print("ok", float(mu.abs().max()), float(var.min()), float(l.sum()), float(var2.max()))
