"""One-off validation at N=32768 (8.6 GB matrix): factor residual against the covariance, agreement with cuSOLVER,
throughput of potrf / solve / a slice of the posterior.  Not part of the test-suite (takes ~20 s of GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
X, Y, th = wl.single_path(N, 9, 2, 1)
torch.cuda.synchronize(); t0 = time.perf_counter()
m = GPmap.fit_gp(X, Y, theta=th); torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m = GPmap.fit_gp(X, Y, theta=th, check=False); e1.record(); torch.cuda.synchronize()
print(f"N={N}: fit {e0.elapsed_time(e1):.1f} ms (potrf flops N^3/3 -> {N**3/3/e0.elapsed_time(e1)/1e9:.2f} TF incl. cov+solve), info={int(m.info.item())}, lml={m.lml}")
# residual on a block sample: (L L^T)[rows, :] vs K[rows, :] for 512 random rows
L = m.K[:, :N]
rows = torch.randperm(N, device="cuda")[:512].sort().values
Lt = torch.tril(L)
prod = Lt[rows] @ Lt.T
from oracle import gp_ref
Kref = torch.from_numpy(gp_ref.cross_cov(X[rows.cpu().numpy()], X, th)).cuda()
Kref[torch.arange(512), rows] += th[3]
print("max |L L^T - K| on 512 rows:", float((prod - Kref).abs().max()))
del prod, Kref, Lt
# K alpha = y on those rows
r = torch.from_numpy(gp_ref.cross_cov(X[rows.cpu().numpy()], X, th)).cuda() @ m.alpha + th[3] * m.alpha[rows] - torch.from_numpy(Y).cuda()[rows]
print("max |K alpha - y| on 512 rows:", float(r.abs().max()))
e0.record(); mu, var = m.predict_grid(wl.BOX, (512, 256)); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"predict 512x256 points: {ms:.1f} ms -> {float(N)**2*512*256/ms/1e9:.2f} TF; var in [{float(var.min()):.3e}, {float(var.max()):.3f}]")
