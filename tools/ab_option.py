"""A/B of one library option on the single-matrix fit (cov + potrf + solves through gpm_fit) and on potrf alone:
python tools/ab_option.py <option> [N ...]   (median of 9 alternating runs each)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gaussianprocesspathmodelling_b200 import _native, workloads as wl
opt = sys.argv[1]
sizes = [int(v) for v in sys.argv[2:]] or [1024, 2048, 4096]
lib = _native.load(); h = _native.handle(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream); p = lambda t: C.c_void_p(t.data_ptr())
for N in sizes:
    X, Y, th = wl.single_path(N, 4, 2, 1)
    Xd = torch.from_numpy(X).cuda(); ld = (N + 15) // 16 * 16
    K = torch.empty((N, ld), dtype=torch.float64, device="cuda")
    ws = torch.empty(int(lib.gpm_potrf_workspace_bytes(N)) // 8, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    tha = _native.theta_array(th)
    def run():
        _native.check(lib.gpm_cov(h, p(Xd), N, 2, tha, p(K), ld, 1, st), "cov")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _native.check(lib.gpm_potrf(h, p(K), N, ld, p(ws), p(info), st), "potrf"); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    t = {0: [], 1: []}; Ls = {}
    for rep in range(11):
        for val in (0, 1):
            with _native.option(opt, val):
                ms = run()
            if rep >= 2: t[val].append(ms)
            Ls[val] = torch.tril(K[:, :N]).clone()
    print(f"N={N}: potrf {opt}=0 {np.median(t[0]):.4f} ms | {opt}=1 {np.median(t[1]):.4f} ms | factors bitwise equal: {bool(torch.equal(Ls[0], Ls[1]))}, info={int(info.item())}")
