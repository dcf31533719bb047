"""Multi-GPU host logic on CPU: shard arithmetic and the gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussianprocesspathmodelling_b200 import dist as gdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 4096, 262144, 4194304):
        for w in (1, 2, 3, 4, 8):
            rs = [gdist.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and sizes == gdist.shard_counts(n, w)


def test_round_robin_covers_once():
    for n, w in ((64, 8), (10, 4), (3, 8)):
        seen = sorted(i for r in range(w) for i in gdist.round_robin(n, r, w))
        assert seen == list(range(n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 3, dtype=torch.float64).view(n, 3)
        lo, hi = gdist.shard_range(n, rank, world)
        got = gdist.all_gather_rows(full[lo:hi].clone(), gdist.shard_counts(n, world))
        ok = torch.equal(got, full)
        # equal-count fast path
        full2 = torch.arange(8, dtype=torch.float64)
        lo, hi = gdist.shard_range(8, rank, world)
        got2 = gdist.all_gather_rows(full2[lo:hi].clone(), gdist.shard_counts(8, world))
        # in place: every rank fills only its slice of the full buffer, the gather completes it (even and uneven)
        ok3 = True
        for m in (8, 7):
            want = torch.arange(m * 2, dtype=torch.float64).view(m, 2)
            buf = torch.full((m, 2), -1.0, dtype=torch.float64)
            lo, hi = gdist.shard_range(m, rank, world)
            buf[lo:hi] = want[lo:hi]
            ok3 = ok3 and torch.equal(gdist.all_gather_inplace(buf, gdist.shard_counts(m, world)), want)
        q.put((rank, bool(ok and ok3 and torch.equal(got2, full2))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_rows_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == [(0, True), (1, True)]
