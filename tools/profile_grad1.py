import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
X, Y, th = wl.single_path(4096, 4, 2, 2)
m = GPmap.fit_gp(X, Y, theta=th)
g = m.lml_grad(); g = m.lml_grad(); print(g[0])
