"""cov + potrf only (config 4 by default), for ncu launch lists."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import _native, workloads as wl  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=16384); ap.add_argument("--reps", type=int, default=2); ap.add_argument("--pad", type=int, default=0)
a = ap.parse_args()
N = a.n
X, Y, th = wl.single_path(N, 4, 2, 1)
lib = _native.load(); h = _native.handle(0)
Xd = torch.from_numpy(X).cuda(); ld = (N + 15) // 16 * 16 + a.pad
K = torch.empty((N, ld), dtype=torch.float64, device="cuda")
ws = torch.empty(int(lib.gpm_potrf_workspace_bytes(N)) // 8, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
tha = _native.theta_array(th)
for r in range(a.reps):
    _native.check(lib.gpm_cov(h, p(Xd), N, 2, tha, p(K), ld, 1, st), "cov")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _native.check(lib.gpm_potrf(h, p(K), N, ld, p(ws), p(info), st), "potrf"); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {r}: potrf N={N} {ms:.3f} ms = {N**3 / 3 / ms / 1e9:.2f} TFLOP/s, info={int(info.item())}")
