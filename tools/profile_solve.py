"""fit at N (default 16384) then repeated gpm_solve_lml calls, for ncu captures of the chained solve."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=16384); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
X, Y, th = wl.single_path(a.n, 4, 2, 2)
m = GPmap.fit_gp(X, Y, theta=th)
lib = _native.load(); h = _native.handle(0)
Yd = torch.from_numpy(Y).cuda(); al = torch.empty_like(Yd); lml = torch.empty(2, dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _native.check(lib.gpm_solve_lml(h, p(m.K), a.n, m.K.stride(0), p(m.ws), p(Yd), 2, p(al), p(lml), st), "solve"); e1.record()
    torch.cuda.synchronize(); print(f"solve N={a.n}: {e0.elapsed_time(e1):.3f} ms")
