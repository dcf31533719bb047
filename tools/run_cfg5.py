"""BASELINE config 5: N=16384 model, posterior mean+variance on a 2048x2048 grid sharded across the
ranks of one node, plus the 64-point LML hyper-parameter sweep (round-robin over ranks).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_cfg5.py [--rows R] [--sweep S]

--rows limits the grid to its first R rows (for short single-GPU runs).  One JSON line on rank 0.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist  # noqa: E402
from gaussianprocesspathmodelling_b200 import GPmap, workloads as wl, dist as gdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384); ap.add_argument("--g", type=int, default=2048)
ap.add_argument("--rows", type=int, default=0); ap.add_argument("--sweep", type=int, default=64)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the VERSION and WARN levels,
    # so those levels are dropped and whatever NCCL does log goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def maxr(x):
    if world == 1: return x
    t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

X, Y, th = wl.single_path(a.n, 5, 2, 2)
Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
Gx, Gy = a.g, (a.rows or a.g)
y1 = wl.BOX[2] + (wl.BOX[3] - wl.BOX[2]) * (Gy - 1) / (a.g - 1) if a.g > 1 else wl.BOX[3]
bounds = (wl.BOX[0], wl.BOX[1], wl.BOX[2], y1)
# warm-up on a sliver, then the timed fit + sharded prediction
m = GPmap.fit_gp(Xd, Yd, theta=th); m.predict_grid(bounds, (Gx, Gy), points=(0, 256)); del m
sync()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
m = GPmap.fit_gp(Xd, Yd, theta=th, check=False)
e1.record()
lo, hi = gdist.shard_range(Gx * Gy, rank, world)
mu, var = m.predict_grid(bounds, (Gx, Gy), points=(lo, hi))
e2.record()
sync()
t_fit, t_pred = maxr(e0.elapsed_time(e1)), maxr(e1.elapsed_time(e2))
t0 = time.perf_counter()
counts = gdist.shard_counts(Gx * Gy, world)
mu_all = gdist.all_gather_rows(mu, counts) if world > 1 else mu
var_all = gdist.all_gather_rows(var, counts) if world > 1 else var
sync()
t_gather = maxr((time.perf_counter() - t0) * 1e3)
ok = bool(bool(torch.isfinite(var_all).all()) and float(var_all.min()) > -1e-8 and float(var_all.max()) <= float(th[2]) + 1e-9)
info = int(m.info.item())
del m, mu, var
torch.cuda.empty_cache()
# hyper-parameter sweep
ths = wl.sweep_thetas(D=2)[: a.sweep]
sync(); t0 = time.perf_counter()
table = gdist.lml_sweep_sharded(Xd, Yd, ths) if a.sweep > 0 else np.zeros((0, 2))
sync(); t_sweep = maxr(time.perf_counter() - t0)
if rank == 0:
    M = Gx * Gy
    print(json.dumps({"config": f"cfg5: N={a.n}, grid {Gx}x{Gy}, sweep S={len(ths)}", "n_gpus": world,
                      "fit_ms": t_fit, "predict_ms": t_pred, "gather_ms": t_gather,
                      "grid_points_per_s": M / ((t_fit + t_pred) * 1e-3), "predict_tflops_total": float(a.n) ** 2 * M / (t_pred * 1e-3) / 1e12,
                      "sweep_s": t_sweep, "sweep_fits_per_s": (len(ths) / t_sweep) if len(ths) else None,
                      "lml_best": [float(v) for v in table.max(axis=0)] if len(ths) else None,
                      "variance_in_bounds": ok, "potrf_info": info}))
if world > 1:
    dist.destroy_process_group()
