#!/usr/bin/env python
"""Benchmark of the GP path-modelling hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extra] [--no-cpu]

A "step" is one pass of the hot path over BASELINE.json config 2 (`configs[1]`): one GP with N=4096 path samples,
fit (covariance -> Cholesky -> alpha + LML) and posterior mean + variance on a 512x512 grid.  metric = posterior
grid points/s.  Under torchrun (N>1 ranks) every rank fits the (replicated, deterministic) model and evaluates its
own 512x512 share of a 512 x (512*N) grid straight into its slice of the full result buffers, which ONE IN-PLACE
NCCL ALL-GATHER per buffer completes on every rank -- inside the timed region (weak scaling; the gather is the only
collective of the path).  `extra` carries the other headline numbers of BASELINE.json's metric, measured in the same
run: batched fits/s (config 3, 4096 paths x N=512 per GPU), Cholesky TFLOP/s at N=16384 (config 4), and config 5
strong-scaled (fixed 2048x2048 grid on the N=16384 model + the 64-point LML sweep; one GPU runs a 1/8 slice).

`--impl reference`: the reference has no GP implementation, so this arm times the
numpy/scipy oracle port of the same path on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import os
import sys


def _usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:                                           # noqa: BLE001
        return os.cpu_count() or 1


def _cpu_arm_bootstrap():
    """The CPU arms (`--impl reference`, `--cpu-sample`) must see the same thread pools whether they are started by
    hand or by torchrun (which exports OMP_NUM_THREADS=1 to every rank): the BLAS / OpenMP pools read their sizes when
    numpy loads, so the environment is fixed and the interpreter re-executed BEFORE numpy is imported."""
    argv = sys.argv[1:]
    is_ref = any(a == "--impl=reference" for a in argv) or \
        any(a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference" for i, a in enumerate(argv))
    if not (is_ref or "--cpu-sample" in argv):
        return
    if is_ref and int(os.environ.get("RANK", "0")) != 0:
        sys.exit(0)                                             # only rank 0 runs the CPU arm
    if os.environ.get("GPM_CPU_ARM_ENV") == "1":
        return
    env = dict(os.environ)
    n = str(_usable_cores())
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "VECLIB_MAXIMUM_THREADS"):
        env[k] = n
    env["GPM_CPU_ARM_ENV"] = "1"
    sys.stdout.flush()
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


_cpu_arm_bootstrap()

import argparse      # noqa: E402
import json          # noqa: E402
import subprocess    # noqa: E402
import threading     # noqa: E402
import time          # noqa: E402

import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "posterior grid points/sec (GP N=4096 fit + mean+variance on a 512x512 grid)"
UNIT = "grid points/s"
CFG = dict(N=4096, D=2, R=2, G=512, seed=2)
FLOP_PER_FIT_N512 = 5.1e7          # N^3/3 + 2 N^2 R + c_exp N^2/2 at N=512, R=2


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback", "fp64_tflops": 37.0, "fp64_src": "nominal"}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"], peaks["hbm_src"] = float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:                                           # noqa: BLE001
        pass
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_FP64.json")))
        peaks["fp64_tflops"], peaks["fp64_src"] = float(p["dgemm_tflops_burst"]), "measured cuBLAS DGEMM 8192^3 (MEASURED_FP64.json)"
        peaks["fp64_tflops_sustained"] = float(p["dgemm_tflops_sustained"])
    except Exception:                                           # noqa: BLE001
        pass
    return peaks


def ncu_traffic():
    """DRAM bytes (read + write) of the dominant launch -- the fused variance sweep of gemm_nt_kernel -- from the
    committed `ncu --set full` capture of the same workload."""
    for name in ("r02_ncu_gemm_summary.json", "r01_final2_ncu_gemm_summary.json"):
        try:
            s = json.load(open(os.path.join(ROOT, "profiles", name)))
            N, M = CFG["N"], CFG["G"] ** 2
            return {"bytes_per_launch": s["dram_read_bytes"] + s["dram_write_bytes"],
                    "launches_per_step": s.get("launches_per_step", 1),
                    "algorithmic_bytes_section_8d": 2.0 * 8.0 * N * M / s.get("launches_per_step", 1),
                    "blocked_algorithm_operand_bytes": s.get("blocked_algorithm_operand_bytes", s.get("algorithmic_operand_bytes")),
                    "note": "all byte figures per launch (a step is launches_per_step chunk launches). The sweep re-reads "
                            "W[:,0:k] for every block column k (blocked forward substitution), so its operand traffic is 8.6x "
                            "the 2*8*N*M bytes of section 8d; at 1.2 TB/s this is far below the HBM roof and the kernel is "
                            "DMMA-issue bound",
                    "duration_ms_under_ncu": s["duration_ms"], "dmma_pipe_active_pct": s["dmma_pipe_active_pct"],
                    "source": "profiles/" + name}
        except Exception:                                       # noqa: BLE001
            continue
    return None


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:                                       # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = max(mx, float(f[1]))
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[0])); power = max(power, float(f[6]))
                    for n, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": power}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of config 2 (+ config 3), per-step times
# ------------------------------------------------------------------------------------------------
def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:                                           # noqa: BLE001
        return _usable_cores()


def cpu_sample(n_query=16384):
    """Oracle (numpy/scipy) on config 2: full fit + posterior on a sample of the grid, extrapolated linearly in the
    number of points to the whole 512x512 grid.  Returns (points_per_s, detail) with per-step wall times."""
    from oracle import gp_ref
    from gaussianprocesspathmodelling_b200 import workloads as wl
    X, Y, th = wl.single_path(CFG["N"], CFG["seed"], CFG["D"], CFG["R"])
    M = CFG["G"] ** 2
    P = gp_ref.grid_points(wl.BOX, (CFG["G"], CFG["G"]))
    idx = np.linspace(0, M - 1, n_query).astype(np.int64)
    m, tf = gp_ref.fit_phases(X, Y, th)
    _, _, tp = gp_ref.predict_phases(m, P[idx])
    t_fit = sum(tf.values())
    t_pred = sum(tp.values())
    scale = M / n_query
    est = t_fit + t_pred * scale
    detail = {"fit_s": t_fit, "predict_sample_s": t_pred, "sample_points": int(n_query), "extrapolated_step_s": est,
              "blas_threads": _blas_threads(), "host_cores": _usable_cores(),
              "per_step_s_full_grid": {"cov": tf["cov_s"], "cholesky": tf["cholesky_s"], "solve_lml": tf["solve_lml_s"],
                                       "cross_cov": tp["cross_cov_s"] * scale, "mean": tp["mean_s"] * scale,
                                       "variance": tp["var_s"] * scale}}
    return M / est, detail


def _fit_one(args):
    from oracle import gp_ref
    return gp_ref.fit(*args)["lml"]


def _pool_init():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:                                           # noqa: BLE001
        pass


def cpu_cfg3_sample(n_serial=48, n_pool_per_core=12):
    """Config 3 (N=512 per-path fits) on the host: the oracle's serial Python loop (BLAS threads = all cores, as the
    reference's style would run) and a multiprocessing pool with one single-threaded worker per core."""
    import multiprocessing as mp
    from oracle import gp_ref
    from gaussianprocesspathmodelling_b200 import workloads as wl
    cores = _usable_cores()
    n_pool = max(cores * n_pool_per_core, 32)
    Xb, Yb, th = wl.batched_paths(max(n_serial, n_pool), 512, seed=3, D=3, R=2)
    gp_ref.fit(Xb[0], Yb[0], th)
    t0 = time.perf_counter()
    gp_ref.fit_batched(Xb[:n_serial], Yb[:n_serial], th)
    t_serial = time.perf_counter() - t0
    with mp.get_context("fork").Pool(cores, initializer=_pool_init) as pool:
        pool.map(_fit_one, [(Xb[i], Yb[i], th) for i in range(cores)])             # warm the workers
        t0 = time.perf_counter()
        pool.map(_fit_one, [(Xb[i], Yb[i], th) for i in range(n_pool)], chunksize=max(1, n_pool // (4 * cores)))
        t_pool = time.perf_counter() - t0
    return {"serial_loop_fits_per_s": n_serial / t_serial, "serial_sample_paths": n_serial,
            "pool_fits_per_s": n_pool / t_pool, "pool_sample_paths": n_pool, "pool_workers": cores,
            "unit": "fits/s", "note": "N=512 D=3 R=2 per-path fit + LML; serial loop uses all BLAS threads per fit, "
                                      "the pool runs one single-threaded fit per core"}


def cpu_kmeans_sample(P=1500, k=8, n=33):
    """One Lloyd iteration of the reference's k-means (the oracle's restatement of GPmap.py:65-121, Python loops over
    paths and centroids with numpy inside calc_distance, as the reference runs it) on a bounded sample of paths."""
    from oracle import gp_ref
    from gaussianprocesspathmodelling_b200 import workloads as wl
    xs, ys, ts = wl.trajectory_families(P, k, n, seed=5)
    t0 = time.perf_counter()
    gp_ref.lloyd(xs, ys, ts, list(range(k)), threshold=0.0, max_iter=1)
    dt = time.perf_counter() - t0
    return {"sample_paths": P, "clusters": k, "s_per_iteration_on_sample": dt, "path_assignments_per_s": P / dt,
            "note": "single-threaded by construction (a Python double loop, GPmap.py:72-80)"}


def run_cpu_sample(args):
    """`--cpu-sample`: print the CPU baseline as one JSON object (run by the B200 arm in a clean subprocess)."""
    cpu_sample(1024)                                            # warm-up (imports, BLAS thread pools)
    v, det = cpu_sample(args.cpu_points)
    out = {"value": v, "detail": det, "cfg3": cpu_cfg3_sample(), "kmeans": cpu_kmeans_sample()}
    print(json.dumps(out))


def cpu_baseline_subprocess(points=16384):
    """Run the CPU baseline in a fresh interpreter with the host's full thread pools (this process may have been
    started by torchrun with OMP_NUM_THREADS=1)."""
    env = {k: v for k, v in os.environ.items() if k != "GPM_CPU_ARM_ENV"}
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample", "--cpu-points", str(points)],
                       capture_output=True, text=True, env=env, timeout=900)
    if r.returncode != 0:
        return {"error": r.stderr[-400:]}
    return json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])


def run_reference(args):
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(2048)
    vals, det = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, det = cpu_sample(8192)
        vals.append(v)
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": CFG["G"] ** 2 / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: N=4096 D=2 R=2, fit + mean+variance on 512x512 grid", "seed": CFG["seed"],
                   "note": "reference has no GP code; numpy/scipy oracle port of the same path on host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": det["blas_threads"], "kind": "port",
                         "sample": f"full N=4096 fit + posterior on {det['sample_points']} of 262144 grid points per step, "
                                   f"extrapolated linearly in the number of points; wall {wall:.1f}s",
                         "per_step_s_full_grid": det["per_step_s_full_grid"], "host_cores": det["host_cores"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def ev_pair(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(torch, fn, reps, warm=1, flush=None):
    """Mean CUDA-event time (ms) of fn on the current stream, after `warm` untimed calls."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush()
        e0, e1 = ev_pair(torch)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def timed_pair(torch, fa, fb, reps, flush=None):
    """Median CUDA-event times (ms) of two functions measured alternately (a, b, a, b, ...): a difference of the two
    is then free of the drift between two separate timing loops."""
    fa(); fb()
    torch.cuda.synchronize()
    ta, tb = [], []
    for _ in range(reps):
        for f, acc in ((fa, ta), (fb, tb)):
            if flush is not None:
                flush()
            e0, e1 = ev_pair(torch)
            e0.record(); f(); e1.record(); e1.synchronize()
            acc.append(e0.elapsed_time(e1))
    return float(np.median(ta)), float(np.median(tb))


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from gaussianprocesspathmodelling_b200 import GPmap, _native, workloads as wl
    from gaussianprocesspathmodelling_b200 import dist as gdist
    from gaussianprocesspathmodelling_b200.dist import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the VERSION and WARN levels,
        # so those levels are dropped and whatever NCCL does log goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _native.load()
    h = _native.handle(local)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N, D, R, G = CFG["N"], CFG["D"], CFG["R"], CFG["G"]
    X, Y, th = wl.single_path(N, CFG["seed"], D, R)
    Xd, Yd = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
    Gy_total = G * world                                 # weak scaling: the grid grows by 512 rows per GPU
    bounds = wl.BOX
    shape = (G, Gy_total)
    M_total = G * Gy_total
    lo, hi = shard_range(M_total, rank, world)
    M_local = hi - lo
    # full-size result buffers: every rank's kernels write their share in place, the all-gather completes them
    mu_full = torch.empty((M_total, R), dtype=torch.float64, device=dev)
    var_full = torch.empty((M_total,), dtype=torch.float64, device=dev)
    state = {}

    def step_device():
        m = GPmap.fit_gp(Xd, Yd, theta=th, check=False)
        gdist.predict_grid_sharded(m, bounds, shape, gather=True, out=(mu_full, var_full))
        state["model"] = m

    # ---- timed region 1: device-resident inputs (public API; the gather is inside the region) --------
    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks = Clocks(local)
    clocks.start()
    time.sleep(0.3)
    launches0 = lib.gpm_launch_count()
    barrier()
    t0w = time.time()
    e0, e1 = ev_pair(torch)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    t1w = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = (lib.gpm_launch_count() - launches0) // args.steps
    clk = clocks.stop(t0w, t1w)
    ms_step = ms_total / args.steps
    value = M_total / (ms_step * 1e-3)
    model = state["model"]
    info = int(model.info.item())

    # ---- timed region 2: end to end from NUMPY inputs (what a GPmap.py user passes) through the public API:
    #      numpy -> cached pinned staging -> H2D -> fit_gp -> sharded predict_grid (+ in-place NCCL all-gather)
    #      -> D2H of the full gathered mean and variance into pinned host buffers, every step ---------------
    mu_h = torch.empty((M_total, R), dtype=torch.float64).pin_memory()
    var_h = torch.empty((M_total,), dtype=torch.float64).pin_memory()

    def step_e2e():
        m = GPmap.fit_gp(X, Y, theta=th, check=False)            # numpy in
        gdist.predict_grid_sharded(m, bounds, shape, gather=True, out=(mu_full, var_full))
        mu_h.copy_(mu_full, non_blocking=True); var_h.copy_(var_full, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return m

    for _ in range(3):
        step_e2e()
    barrier()
    e0, e1 = ev_pair(torch)
    e2e_wall = []
    e0.record()
    for _ in range(args.steps):
        tw = time.perf_counter()
        step_e2e()
        e2e_wall.append((time.perf_counter() - tw) * 1e3)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = M_total / (ms_e2e * 1e-3)
    h2d = X.nbytes + Y.nbytes
    d2h = mu_h.numel() * 8 + var_h.numel() * 8

    # the collective alone (events around the two in-place all-gathers of one step)
    gather_ms = 0.0
    if world > 1:
        counts = gdist.shard_counts(M_total, world)

        def only_gather():
            gdist.all_gather_inplace(mu_full, counts); gdist.all_gather_inplace(var_full, counts)
        gather_ms = max_over_ranks(timed(torch, only_gather, 5, warm=2))

    # ---- per-kernel phases on this rank (CUDA events on the launching stream) ---------------------
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)   # noqa: E731
    tha = _native.theta_array(th)
    ld = (N + 15) // 16 * 16
    K = torch.empty((N, ld), dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.gpm_potrf_workspace_bytes(N)) // 8, dtype=torch.float64, device=dev)
    infod = torch.zeros(1, dtype=torch.int32, device=dev)
    alpha = torch.empty((N, R), dtype=torch.float64, device=dev)
    lml = torch.empty((R,), dtype=torch.float64, device=dev)
    mu, var = mu_full[lo:hi], var_full[lo:hi]
    grid = _native.GpmGrid(bounds[0], bounds[1], bounds[2], bounds[3], 0.0, G, Gy_total)
    pws_bytes = min(int(lib.gpm_predict_workspace_bytes(h, N, M_local)), GPmap.PREDICT_WORKSPACE_BYTES)
    pws = GPmap._workspace(pws_bytes, dev, "predict")          # the cached buffer the timed loop used
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flushbuf.zero_()

    def k_cov(flags=0):
        _native.check(lib.gpm_cov(h, ptr(Xd), N, D, tha, ptr(K), ld, flags, st), "cov")

    def k_potrf():
        _native.check(lib.gpm_potrf(h, ptr(K), N, ld, ptr(ws), ptr(infod), st), "potrf")

    def k_solve():
        _native.check(lib.gpm_solve_lml(h, ptr(K), N, ld, ptr(ws), ptr(Yd), R, ptr(alpha), ptr(lml), st), "solve")

    def k_pred(flags):
        _native.check(lib.gpm_predict(h, ptr(Xd), N, D, tha, ptr(K), ld, ptr(ws), ptr(alpha), R, None, C.byref(grid),
                                      lo, hi, ptr(mu), ptr(var), ptr(pws), pws.numel() * 8, flags, st), "predict")

    phases = {}
    phases["cov_full_ms"] = timed(torch, lambda: k_cov(0), 5, flush=flush)
    phases["cov_lower_ms"] = timed(torch, lambda: k_cov(1), 5, flush=flush)

    def cov_potrf():
        k_cov(1); k_potrf()
    t_cp = timed(torch, cov_potrf, 5, flush=flush)
    phases["potrf_ms"] = t_cp - phases["cov_lower_ms"]
    phases["solve_lml_3step_ms"] = timed(torch, k_solve, 5, flush=flush)     # gpm_solve_lml alone: forward + backward chain + LML

    def k_fit():
        _native.check(lib.gpm_fit(h, ptr(Xd), N, D, tha, ptr(Yd), R, ptr(K), ld, ptr(ws), ptr(alpha), ptr(lml), ptr(infod), st), "fit")
    # the solve's share of the fused fit (forward substitution rides on the factorisation; the backward chain, which
    # also produces the LML, after it): gpm_fit against covariance + factorisation, measured alternately, medians of 15
    t_cp_i, t_fit_i = timed_pair(torch, cov_potrf, k_fit, 15, flush=flush)
    phases["fit_fused_ms"] = t_fit_i                                         # gpm_fit: what GPmap.fit_gp issues
    phases["solve_lml_ms"] = t_fit_i - t_cp_i
    nl0 = lib.gpm_launch_count()
    phases["predict_var_ms"] = timed(torch, lambda: k_pred(2), 2, warm=1)
    var_launches = (lib.gpm_launch_count() - nl0) // 3
    phases["predict_mean_ms"] = timed(torch, lambda: k_pred(1), 3, warm=1)
    # materialised cross-covariance K*^T of this rank's grid share into the variance workspace (separable grid
    # kernel: a pure HBM write stream of 8 N M bytes)
    npad_cc = (N + 127) // 128 * 128
    cc_rows = min(M_local, (pws.numel() * 8) // (npad_cc * 8 + 8))

    def k_cross():
        _native.check(lib.gpm_cross_cov(h, ptr(Xd), N, D, tha, None, C.byref(grid), lo, lo + cc_rows, ptr(pws), npad_cc, st),
                      "cross_cov")
    phases["cross_cov_ms"] = timed(torch, k_cross, 3, warm=1)
    nblk = (N + 127) // 128
    chunks = max(1, var_launches // 3)                       # per chunk: cross-cov + one persistent sweep launch + finalize
    gemm_launches = max(1, var_launches - 2 * chunks)
    var_flops = float(N) * N * M_local
    var_tflops = var_flops / (phases["predict_var_ms"] * 1e-3) / 1e12

    # cuBLAS DGEMM 8192^3 in this very run (MEASURED_PEAKS.json has no FP64 figure; MEASURED_FP64.json is ours)
    peak_same_run = None
    try:
        A = torch.randn(8192, 8192, dtype=torch.float64, device=dev); Bm = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        torch.matmul(A, Bm)
        best = min(timed(torch, lambda: torch.matmul(A, Bm), 1, warm=0) for _ in range(5))
        peak_same_run = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del A, Bm
    except Exception:                                           # noqa: BLE001
        pass
    roofline = {
        "kernel": "gemm_nt_kernel (DMMA.8x8x4 + TMA): the fused blocked-TRSM sweep of the posterior variance, one persistent launch per chunk",
        "bound": "tensor", "achieved": var_tflops, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
        "frac": var_tflops / peaks["fp64_tflops"],
        "peak_same_run": peak_same_run, "frac_of_peak_same_run": (var_tflops / peak_same_run) if peak_same_run else None,
        "traffic": (ncu_traffic() or {}).get("bytes_per_launch"), "traffic_detail": ncu_traffic(),
        "peak_source": peaks["fp64_src"],
        "launches_per_step": int(gemm_launches), "flops_per_launch": var_flops / gemm_launches,
        "avg_launch_ms": phases["predict_var_ms"] / gemm_launches,
        "note": "algorithmic flops N^2*M of V = L^-1 K*; time = CUDA events around the variance phase "
                "(cross-cov write + sweep launches + finalize); peak_same_run = torch.matmul f64 8192^3 best of 5 in this process",
    }
    kernels = {
        "cov_full": {"bound": "hbm", "achieved": 8.0 * N * N / (phases["cov_full_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"},
        "potrf": {"bound": "tensor", "achieved": N ** 3 / 3 / (phases["potrf_ms"] * 1e-3) / 1e12, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s"},
        "solve_lml": {"bound": "hbm", "achieved": 4.0 * N * N / (max(phases["solve_lml_ms"], 1e-6) * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"},
        "cross_cov": {"bound": "hbm", "achieved": 8.0 * N * cc_rows / (phases["cross_cov_ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"},
        "predict_mean": {"bound": "fp64 (separable grid form: 2 shared loads + multiply + R FMAs per pair, 1/8 exp per pair)",
                         "achieved": float(N) * M_local / (phases["predict_mean_ms"] * 1e-3) / 1e9, "peak": None, "unit": "G kernel values/s"},
    }
    for kk in kernels.values():
        kk["frac"] = (kk["achieved"] / kk["peak"]) if kk["peak"] else None
    kernels["cov_full"]["note"] = "134 MB in ~35 us: launch ramp dominates; the N=16384 figure is extra.cfg4_N16384.cov_frac_hbm"
    kernels["potrf"]["note"] = "32 block columns: panel-chain bound; the N=16384 figure is extra.cfg4_N16384.potrf_frac_dgemm"
    kernels["solve_lml"]["note"] = ("the solve's share of gpm_fit (backward chain + LML; the forward pass rides on the factorisation): one pass over "
                                    "the lower triangle of L = 4 N^2 bytes; latency chain of block hand-offs; N=16384: extra.cfg4_N16384.solve_gbs")

    extra = {"phases_ms": phases, "kernels": kernels, "potrf_info": info, "gather_ms": gather_ms}

    # ---- config 1 (the reference-sized case): end-to-end latency, launch-bound, no roofline claim ----
    X1, Y1, th1 = wl.single_path(200, 1, 2, 2)
    X1d, Y1d = torch.from_numpy(X1).to(dev), torch.from_numpy(Y1).to(dev)

    def cfg1():
        m1 = GPmap.fit_gp(X1d, Y1d, theta=th1, check=False)
        return m1.predict_grid(wl.BOX, (100, 100))
    extra["cfg1_N200_100x100_latency_ms"] = timed(torch, cfg1, 20, warm=3)
    try:
        mu_e, var_e = cfg1()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                cfg1()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            mu_g, var_g = cfg1()
        graph.replay(); torch.cuda.synchronize()
        same = bool(torch.equal(mu_g, mu_e) and torch.equal(var_g, var_e))
        extra["cfg1_cuda_graph_latency_ms"] = timed(torch, graph.replay, 50, warm=3)
        extra["cfg1_cuda_graph_matches_eager"] = same
        del graph
    except Exception as e:                                      # noqa: BLE001
        extra["cfg1_cuda_graph_error"] = str(e)[:200]
        torch.cuda.set_stream(torch.cuda.default_stream())      # a failed capture leaves torch on the capture stream

    # ---- the other headline numbers ------------------------------------------------------------------
    if not args.no_extra:
        del K, ws, mu, var, pws
        state.clear(); model = None
        mu_full = var_full = None
        GPmap.clear_workspaces()
        torch.cuda.empty_cache()
        # -- config 4: N=16384 covariance, Cholesky, solve --
        N4 = 16384
        X4, Y4, th4 = wl.single_path(N4, 4, 2, 1)
        X4d = torch.from_numpy(X4).to(dev)
        K4 = torch.empty((N4, N4), dtype=torch.float64, device=dev)
        ws4 = torch.empty(int(lib.gpm_potrf_workspace_bytes(N4)) // 8, dtype=torch.float64, device=dev)
        th4a = _native.theta_array(th4)

        def cov4(flags):
            _native.check(lib.gpm_cov(h, ptr(X4d), N4, 2, th4a, ptr(K4), N4, flags, st), "cov")

        def cp4():
            cov4(1)
            _native.check(lib.gpm_potrf(h, ptr(K4), N4, N4, ptr(ws4), ptr(infod), st), "potrf")
        t_cov4 = timed(torch, lambda: cov4(0), 5, flush=flush)
        t_cov4l = timed(torch, lambda: cov4(1), 5, flush=flush)
        t_cp4 = timed(torch, cp4, 3, flush=flush)
        t_potrf4 = t_cp4 - t_cov4l
        y4 = torch.from_numpy(Y4).to(dev); a4 = torch.empty_like(y4); l4 = torch.empty(1, dtype=torch.float64, device=dev)
        t_solve4 = timed(torch, lambda: _native.check(lib.gpm_solve_lml(h, ptr(K4), N4, N4, ptr(ws4), ptr(y4), 1, ptr(a4), ptr(l4), st), "solve"), 3, flush=flush)
        t_fit4 = timed(torch, lambda: GPmap.fit_gp(X4d, y4, theta=th4, check=False), 3, flush=flush)
        t_cp4_i, t_fused4 = timed_pair(torch, cp4, lambda: _native.check(lib.gpm_fit(h, ptr(X4d), N4, 2, th4a, ptr(y4), 1, ptr(K4), N4, ptr(ws4), ptr(a4), ptr(l4), ptr(infod), st), "fit"), 9, flush=flush)
        extra["cfg4_N16384"] = {
            "cov_ms": t_cov4, "cov_gbs": 8.0 * N4 * N4 / (t_cov4 * 1e-3) / 1e9, "cov_frac_hbm": 8.0 * N4 * N4 / (t_cov4 * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "cov_lower_ms": t_cov4l,
            "potrf_ms": t_potrf4, "potrf_tflops": N4 ** 3 / 3 / (t_potrf4 * 1e-3) / 1e12,
            "potrf_frac_dgemm": N4 ** 3 / 3 / (t_potrf4 * 1e-3) / 1e12 / peaks["fp64_tflops"],
            "solve_lml_3step_ms": t_solve4, "fit_fused_ms": t_fused4,
            # the solve's share of the fused fit: gpm_fit minus (covariance + factorisation), measured alternately; its algorithmic traffic
            # is one pass over the lower triangle of L (4 N^2 bytes) since the forward pass rides on the factorisation
            "solve_lml_ms": t_fused4 - t_cp4_i, "solve_gbs": 4.0 * N4 * N4 / (max(t_fused4 - t_cp4_i, 1e-6) * 1e-3) / 1e9,
            "solve_frac_hbm": 4.0 * N4 * N4 / (max(t_fused4 - t_cp4_i, 1e-6) * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "fit_gp_total_ms": t_fit4,
            "info": int(infod.item()),
        }
        del K4, ws4
        torch.cuda.empty_cache()
        # -- config 3: batched N=512 fits, 4096 paths per GPU, alpha / lml all-gathered in place at N>1 --
        B = 4096
        Xb, Yb, thb = wl.batched_paths(B, 512, seed=3, D=3, R=2, first=rank * B)
        Xbd, Ybd = torch.from_numpy(Xb).to(dev), torch.from_numpy(Yb).to(dev)
        t_b = max_over_ranks(timed(torch, lambda: GPmap.fit_gp_batched(Xbd, Ybd, theta=thb, check=False), 3, warm=2))
        cfg3 = {"paths_per_gpu": B, "N": 512, "ms": t_b, "fits_per_s": world * B / (t_b * 1e-3),
                "frac_of_fp64_ceiling": (B / (t_b * 1e-3)) * FLOP_PER_FIT_N512 / (peaks["fp64_tflops"] * 1e12),
                "ceiling_fits_per_s_per_gpu": peaks["fp64_tflops"] * 1e12 / FLOP_PER_FIT_N512}
        if world > 1:
            counts3 = [B] * world
            t_bg = max_over_ranks(timed(torch, lambda: gdist.fit_gp_batched_sharded(Xbd, Ybd, counts3, theta=thb, check=False), 3, warm=1))
            cfg3["with_gather_ms"] = t_bg
            cfg3["fits_per_s_with_gather"] = world * B / (t_bg * 1e-3)
        extra["cfg3_batched"] = cfg3
        del Xbd, Ybd
        # the reference's own path length (GPmap.py:189 resamples every trajectory to 33 points): one CTA per path
        Bs = 16384
        Xs_, Ys_, ths_ = wl.batched_paths(Bs, 33, seed=3, D=2, R=2, first=rank * Bs)
        Xsd, Ysd = torch.from_numpy(Xs_).to(dev), torch.from_numpy(Ys_).to(dev)
        t_s = max_over_ranks(timed(torch, lambda: GPmap.fit_gp_batched(Xsd, Ysd, theta=ths_, check=False), 5, warm=2))
        extra["short_paths_N33"] = {"paths_per_gpu": Bs, "N": 33, "ms": t_s, "fits_per_s": world * Bs / (t_s * 1e-3),
                                    "kernel": "fit_small_kernel: one CTA per path, whole fit in shared memory"}
        del Xsd, Ysd
        # the reference's real hot loop, trajectories.kmeansclustering (GPmap.py:36-121), device-resident: time per Lloyd
        # iteration (assignment + member-order centroid means + convergence sum) on 100000 synthetic 33-point trajectories
        Pk, kk, nk = 100000, 8, 33
        kx, ky, kt = wl.trajectory_families(Pk, kk, nk, seed=5 + rank)
        kdev = torch.from_numpy(np.stack([kx, ky, kt])).to(dev)
        kxT, kyT = kdev[0].t().contiguous(), kdev[1].t().contiguous()
        ksel = torch.arange(kk, device=dev)
        kws = torch.empty(int(lib.gpm_kmeans_workspace_bytes(Pk, nk, kk)) // 8, dtype=torch.float64, device=dev)
        kassign = torch.zeros(Pk, dtype=torch.int32, device=dev)

        def k_lloyd(iters=10):
            cents = kdev[:, ksel, :].contiguous()
            # threshold 0: never converges early, every enqueued iteration does its full work
            _native.check(lib.gpm_kmeans_lloyd(h, ptr(kdev[0]), ptr(kdev[1]), ptr(kdev[2]), ptr(kxT), ptr(kyT), Pk, nk, kk,
                                               ptr(cents), ptr(kassign), 0.0, iters, 1, ptr(kws), st), "kmeans_lloyd")
        t_k = max_over_ranks(timed(torch, k_lloyd, 3, warm=1)) / 10
        extra["kmeans_lloyd"] = {"paths_per_gpu": Pk, "samples_per_path": nk, "clusters": kk, "ms_per_iteration": t_k,
                                 "path_assignments_per_s": world * Pk / (t_k * 1e-3),
                                 "note": "assignment (calc_distance, GPmap.py:114-121) + centroid means in member order "
                                         "(calc_mean_traj, bit-exact with the reference) + convergence sum, all on the device"}
        del kdev, kxT, kyT, kws
        GPmap.clear_workspaces()
        torch.cuda.empty_cache()
        # -- config 5, STRONG scaling: fixed 2048x2048 grid on the N=16384 model, grid points sharded over the ranks,
        #    in-place all-gather; plus the 64-point LML sweep, round-robin over the ranks.  One GPU runs a 1/8 slice. --
        N5, G5 = 16384, 2048
        X5, Y5, th5 = wl.single_path(N5, 5, 2, 2)
        X5d, Y5d = torch.from_numpy(X5).to(dev), torch.from_numpy(Y5).to(dev)
        GPmap.fit_gp(X5d, Y5d, theta=th5, check=False)                 # warm-up
        barrier()
        ef0, ef1 = ev_pair(torch)
        ef0.record()
        m5 = GPmap.fit_gp(X5d, Y5d, theta=th5, check=False)
        ef1.record(); ef1.synchronize()
        fit5_ms = max_over_ranks(ef0.elapsed_time(ef1))
        M5 = G5 * G5
        m5.predict_grid(wl.BOX, (G5, G5), points=(0, 148 * 128))       # warm-up: kernels + the 4 GiB workspace
        barrier()
        if world == 1:
            sl = M5 // 8
            ep0, ep1 = ev_pair(torch)
            ep0.record()
            m5.predict_grid(wl.BOX, (G5, G5), points=(0, sl))
            ep1.record(); ep1.synchronize()
            pred_ms = ep0.elapsed_time(ep1)
            cfg5 = {"n_gpus": 1, "slice": "1/8 of the 2048x2048 grid (points 0..524287) on one GPU; a full grid is 8x this",
                    "fit_ms": fit5_ms, "predict_slice_ms": pred_ms, "predict_full_grid_ms_extrapolated": pred_ms * 8,
                    "gather_ms": 0.0, "points_per_s": sl / (pred_ms * 1e-3),
                    "tflops": float(N5) * N5 * sl / (pred_ms * 1e-3) / 1e12}
        else:
            mu5 = torch.empty((M5, 2), dtype=torch.float64, device=dev)
            var5 = torch.empty((M5,), dtype=torch.float64, device=dev)
            tm = {}
            barrier()
            gdist.predict_grid_sharded(m5, wl.BOX, (G5, G5), gather=True, out=(mu5, var5), timings=tm)
            torch.cuda.synchronize()
            ev = tm["events"]
            comp_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
            total_ms = max_over_ranks(ev[0].elapsed_time(ev[2]))
            gat_ms = max_over_ranks(ev[1].elapsed_time(ev[2]))
            cfg5 = {"n_gpus": world, "fit_ms": fit5_ms, "predict_compute_ms": comp_ms, "gather_ms": gat_ms,
                    "predict_total_ms": total_ms, "points_per_s": M5 / (total_ms * 1e-3),
                    "points_per_s_incl_fit": M5 / ((total_ms + fit5_ms) * 1e-3),
                    "tflops_aggregate": float(N5) * N5 * M5 / (total_ms * 1e-3) / 1e12,
                    "note": "strong scaling: fixed 2048x2048 grid; the fit (replicated on every rank) is the Amdahl term, "
                            "gather_ms = the two in-place NCCL all-gathers (max over ranks, includes waiting for the slowest rank)"}
            del mu5, var5
        ths5 = wl.sweep_thetas(D=2)
        if world == 1:
            ths5 = ths5[::8]
        barrier()
        es0, es1 = ev_pair(torch)
        es0.record()
        table = gdist.lml_sweep_sharded(X5d, Y5d, ths5)
        es1.record(); es1.synchronize()
        sw_ms = max_over_ranks(es0.elapsed_time(es1))
        cfg5["sweep"] = {"points": int(len(ths5)), "ms": sw_ms, "fits_per_s": len(ths5) / (sw_ms * 1e-3),
                         "finite": bool(np.isfinite(table).all()),
                         "note": "8 of the 64 points on one GPU" if world == 1 else "64 points round-robin over the ranks, all-reduce of the table"}
        extra["cfg5_strong"] = cfg5
        del m5

    if world > 1:
        dist.destroy_process_group()

    # ---- CPU baseline on this box's host cores (rank 0, every N): a fresh interpreter with full thread pools ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        res = cpu_baseline_subprocess(16384)
        if "error" in res:
            cpu = {"value": None, "unit": UNIT, "cores": _usable_cores(), "kind": "port", "sample": "failed: " + res["error"]}
        else:
            det = res["detail"]
            cpu = {"value": res["value"], "unit": UNIT, "cores": det["blas_threads"], "kind": "port",
                   "sample": f"numpy/scipy oracle: full N=4096 fit ({det['fit_s']:.2f}s) + posterior on 16384 of 262144 grid points "
                             f"({det['predict_sample_s']:.2f}s), extrapolated linearly to the whole grid",
                   "host_cores": det["host_cores"], "per_step_s_full_grid": det["per_step_s_full_grid"],
                   "cfg3_N512_fits": res["cfg3"], "kmeans_lloyd": res.get("kmeans")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg2: GP N={N} D={D} R={R} (seed {CFG['seed']}), fit + posterior mean+variance on a "
                                   f"{G}x{G} grid per GPU ({G}x{Gy_total} total)",
                       "l2": "working set (K 134 MB, W 4 GiB per chunk) exceeds the 126 MB L2; per-kernel phases flush L2 with a 256 MB write",
                       "parallelism": f"grid points sharded over {world} rank(s), model replicated; results written in place into the "
                                      f"full buffers and completed by one in-place NCCL all-gather per buffer inside the timed region"
                       if world > 1 else "single GPU",
                       "e2e_inputs": "numpy arrays through GPmap.fit_gp (cached pinned staging + H2D), D2H of the full mean+variance"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "rank0_step_wall_ms": [round(x, 3) for x in e2e_wall]},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "extra": extra,
        }
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the config-3 / config-4 / config-5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--cpu-sample", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-points", type=int, default=16384, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_sample:
        run_cpu_sample(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
