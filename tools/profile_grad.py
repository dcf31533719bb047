import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
for N in (4096, 16384):
    X, Y, th = wl.single_path(N, 4, 2, 2)
    m = GPmap.fit_gp(X, Y, theta=th)
    m.lml_grad(); torch.cuda.synchronize()
    t0 = time.perf_counter(); g = m.lml_grad(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"N={N}: lml_grad {dt*1e3:.2f} ms ({2*N**3/3/dt/1e12:.2f} TF on 2N^3/3), grad[0]={g[0]}")
