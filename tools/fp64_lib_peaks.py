"""Library FP64 denominators on this B200: cuBLAS DGEMM (torch.matmul f64) burst + sustained, cuSOLVER potrf.
Measurement tool only (writes MEASURED_FP64.json content to stdout); not on the product path."""
import json, time, torch
dev = torch.device("cuda:0")
out = {}
def ev_time(f, reps):
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
for _ in range(3): torch.matmul(a, b)
torch.cuda.synchronize()
ms = ev_time(lambda: torch.matmul(a, b), 10)
out["dgemm_8192_burst_tflops"] = 2 * n**3 / ms * 1e-9
# sustained: back-to-back for ~4 s
t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(10): torch.matmul(a, b)
    cnt += 10
    torch.cuda.synchronize()
e1.record(); e1.synchronize()
out["dgemm_8192_sustained_tflops"] = 2 * n**3 * cnt / e0.elapsed_time(e1) * 1e-9
# NT form (syrk-like): A @ A.T
ms = ev_time(lambda: torch.matmul(a, a.t()), 5)
out["dgemm_8192_nt_tflops"] = 2 * n**3 / ms * 1e-9
del b
for N in (4096, 16384):
    x = torch.randn(N, 64, dtype=torch.float64, device=dev)
    K = x @ x.t() / 64 + torch.eye(N, dtype=torch.float64, device=dev) * 2.0
    torch.linalg.cholesky(K); torch.cuda.synchronize()
    ms = ev_time(lambda: torch.linalg.cholesky(K), 3)
    out[f"cusolver_potrf_{N}_ms"] = ms
    out[f"cusolver_potrf_{N}_tflops"] = N**3 / 3 / ms * 1e-9
    del K, x
print(json.dumps(out))
