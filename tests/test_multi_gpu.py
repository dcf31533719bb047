"""Multi-GPU parity (NCCL, 2 ranks on one box): the sharded entry points of gaussianprocesspathmodelling_b200.dist
must reproduce the single-GPU result BITWISE -- a query's value does not depend on the shard it falls into, paths and
sweep points are independent, and the in-place all-gather only moves bytes.  Skipped with fewer than 2 GPUs
(run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from gaussianprocesspathmodelling_b200 import GPmap, dist as gdist, workloads as wl
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = {}
    try:
        # --- grid prediction: uneven split (101 x 67 points over 2 ranks) and an even one, in-place gather ---
        X, Y, th = wl.single_path(700, seed=21, D=2, R=2)
        m = GPmap.fit_gp(X, Y, theta=th)
        for shape in ((101, 67), (64, 64)):
            mu, var = gdist.predict_grid_sharded(m, wl.BOX, shape, gather=True)
            mu1, var1 = m.predict_grid(wl.BOX, shape)
            res[f"grid{shape}"] = bool(torch.equal(mu, mu1) and torch.equal(var, var1))
        # --- batched fits: 7 paths split 4 + 3 ---
        B, N = 7, 300
        counts = gdist.shard_counts(B, world)
        lo, hi = gdist.shard_range(B, rank, world)
        Xb, Yb, thb = wl.batched_paths(B, N, seed=22, D=3, R=2)
        a, l = gdist.fit_gp_batched_sharded(Xb[lo:hi], Yb[lo:hi], counts, theta=thb)
        a1, l1 = GPmap.fit_gp_batched(Xb, Yb, theta=thb)
        res["batched"] = bool(torch.equal(a, a1) and torch.equal(l, l1))
        # --- hyper-parameter sweep: 6 points round-robin ---
        ths = wl.sweep_thetas(D=2)[::11]
        t = gdist.lml_sweep_sharded(X, Y, ths)
        t1 = GPmap.lml_sweep(X, Y, ths)
        res["sweep"] = bool(np.array_equal(t, t1))
        torch.cuda.synchronize()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_entry_points_equal_single_gpu_bitwise_under_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=500) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(60)
    for rank, res in got:
        assert all(res.values()), (rank, res)
    assert [r for r, _ in got] == [0, 1]
