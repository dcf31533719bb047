import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
N = 16384
X, Y, th = wl.single_path(N, 4, 2, 2)
m = GPmap.fit_gp(X, Y, theta=th); m = GPmap.fit_gp(X, Y, theta=th)
torch.cuda.synchronize()
lib = _native.load()
buf = (C.c_ulonglong * (10 * 8192))()
lib.gpm_debug_solve_ts(buf)
a = np.frombuffer(buf, dtype=np.uint64).reshape(10, 8192)[:, :128].astype(np.int64)
ready, seen, pub = a[0], a[1], a[2]
names = ['z staged', 'tile product', 'y formed', 'inv product', 'z stored', 'published']
seq = [a[3], a[4], a[5], a[6], a[7], a[2]]
prev = seen
for nm, cur in zip(names, seq):
    print('%-14s +%.2f us' % (nm, ((cur - prev)[1:128]).mean() / 1e3)); prev = cur
t0 = pub[0]
print("block: ready-for-last-dep, seen-last-dep, published (us since block 0 published)")
for i in (1, 2, 3, 10, 32, 64, 100, 126, 127):
    print(i, (ready[i] - t0) / 1e3, (seen[i] - t0) / 1e3, (pub[i] - t0) / 1e3)
link = np.diff(pub[:128]) / 1e3
prop = (seen[1:128] - pub[0:127]) / 1e3
proc = (pub[1:128] - seen[1:128]) / 1e3
late = (ready[1:128] - pub[0:127]) / 1e3
print("link us: mean %.2f  | flag propagation (seen[i]-pub[i-1]): mean %.2f  | processing (pub[i]-seen[i]): mean %.2f | ready after pub (bulk lateness): mean %.2f max %.2f" % (link.mean(), prop.mean(), proc.mean(), late.mean(), late.max()))
