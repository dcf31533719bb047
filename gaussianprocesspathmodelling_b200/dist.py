"""Multi-GPU sharding: one process per GPU (torchrun), torch.distributed for the plumbing.

The path shards without any data-path collective: independent paths, grid rows
and hyper-parameter points are split across ranks; the single-matrix Cholesky is replicated on every
rank (deterministic, so every rank holds the same factor -- no broadcast needed).  The only collective
is the final all-gather of the results (NCCL over NVLink on GPUs; gloo in the CPU tests of the index logic).

The gather is IN PLACE: every rank allocates the full result buffer once, its kernels write their share straight
into the rank's slice of it (``out=`` of ``GPModel.predict_grid`` / ``fit_gp_batched``), and the all-gather fills
in the other slices.  No staging copy, no padding, no concatenation.  ``gpm_dist_*`` of
SURVEY.md section 8b is deliberately not a C-ABI group: the collectives are two torch.distributed calls on buffers
the library's kernels have already written, there is nothing for native code to add.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) share of n units for `rank` of `world` (first n % world ranks get one more)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def round_robin(n, rank, world):
    """Indices rank, rank+world, ... < n (used for the hyper-parameter sweep: neighbouring points cost the same)."""
    return list(range(rank, n, world))


def all_gather_inplace(full, counts, group=None):
    """In-place all-gather along dim 0: rank r has already written rows [off_r, off_r + counts[r]) of ``full``;
    on return every rank holds all rows.  Equal counts: one ``all_gather_into_tensor`` whose input is the rank's
    own slice of the output (NCCL's in-place form).  Unequal counts: ``all_gather`` into a list of views of
    ``full`` (NCCL lowers it to one grouped broadcast per rank).  gloo (CPU tests) needs a private copy of the
    input and equal sizes, so unequal counts are padded there -- the NCCL path never pads."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return full
    rank = dist.get_rank(group)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    mine = full[offs[rank]: offs[rank + 1]]
    nccl = full.is_cuda
    if len(set(counts)) == 1:
        dist.all_gather_into_tensor(full, mine if nccl else mine.clone(), group=group)
        return full
    views = [full[offs[r]: offs[r + 1]] for r in range(world)]
    if nccl:
        dist.all_gather(views, mine, group=group)
        return full
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tuple(full.shape[1:]), dtype=full.dtype)
    padded[: counts[rank]] = mine
    flat = torch.empty((world * cmax,) + tuple(full.shape[1:]), dtype=full.dtype)
    dist.all_gather_into_tensor(flat, padded, group=group)
    for r in range(world):
        views[r].copy_(flat[r * cmax: r * cmax + counts[r]])
    return full


def all_gather_rows(local, counts, group=None):
    """All-gather tensors that differ only in their first dimension (counts[r] rows on rank r): allocates the full
    buffer, places ``local`` in this rank's slice and gathers in place.  Prefer writing into the slice directly."""
    import torch
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return local
    rank = dist.get_rank(group)
    full = torch.empty((sum(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    off = sum(counts[:rank])
    full[off: off + counts[rank]] = local
    return all_gather_inplace(full, counts, group)


def _rank_world(group):
    import torch.distributed as dist
    return (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)


def predict_grid_sharded(model, bounds, shape, t=None, return_var=True, include_noise=False, group=None,
                         gather=True, out=None, timings=None):
    """Each rank evaluates its contiguous share of the Gx*Gy grid points straight into its slice of the full
    result buffers; one in-place all-gather per buffer completes them on every rank.

    Returns mu (Gy, Gx, R)[, var (Gy, Gx)] (full grids when ``gather``; otherwise only this rank's slice is valid).
    ``out=(mu_flat, var_flat)`` reuses caller-owned full-size flat buffers.  ``timings`` (dict) receives CUDA events
    bracketing the compute and the gather."""
    import torch
    rank, world = _rank_world(group)
    Gx, Gy = int(shape[0]), int(shape[1])
    M = Gx * Gy
    R = model.alpha.shape[1]
    lo, hi = shard_range(M, rank, world)
    dev = model.X.device
    if out is not None:
        mu_full, var_full = out if return_var else (out, None)
    else:
        mu_full = torch.empty((M, R), dtype=torch.float64, device=dev)
        var_full = torch.empty((M,), dtype=torch.float64, device=dev) if return_var else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timings is not None else None
    if ev:
        ev[0].record()
    local_out = (mu_full[lo:hi], var_full[lo:hi]) if return_var else mu_full[lo:hi]
    model.predict_grid(bounds, shape, t=t, return_var=return_var, include_noise=include_noise, points=(lo, hi),
                       out=local_out)
    if ev:
        ev[1].record()
    if gather and world > 1:
        counts = shard_counts(M, world)
        all_gather_inplace(mu_full, counts, group)
        if return_var:
            all_gather_inplace(var_full, counts, group)
    if ev:
        ev[2].record()
        timings["events"] = ev
    if return_var:
        return mu_full.view(Gy, Gx, R), var_full.view(Gy, Gx)
    return mu_full.view(Gy, Gx, R)


def fit_gp_batched_sharded(Xb_local, Yb_local, counts, group=None, gather=True, **kw):
    """Each rank fits its own paths (already local) into its slice of the full alpha / lml buffers, which one
    in-place all-gather each completes in rank order."""
    import torch
    from .GPmap import fit_gp_batched, _dev
    rank, world = _rank_world(group)
    if not gather or world == 1:
        return fit_gp_batched(Xb_local, Yb_local, **kw)
    Xb_local = _dev(Xb_local)
    Yb_local = _dev(Yb_local, Xb_local.device)
    if Yb_local.ndim == 2:
        Yb_local = Yb_local[:, :, None].contiguous()
    _, N, _ = Xb_local.shape
    R = Yb_local.shape[2]
    total = int(sum(counts))
    off = int(sum(counts[:rank]))
    alpha = torch.empty((total, N, R), dtype=torch.float64, device=Xb_local.device)
    lml = torch.empty((total, R), dtype=torch.float64, device=Xb_local.device)
    fit_gp_batched(Xb_local, Yb_local, out=(alpha[off: off + counts[rank]], lml[off: off + counts[rank]]), **kw)
    all_gather_inplace(alpha, counts, group)
    all_gather_inplace(lml, counts, group)
    return alpha, lml


def lml_sweep_sharded(X, Y, thetas, group=None):
    """Round-robin the hyper-parameter points over ranks; the S x R table is completed by one all-reduce over
    disjoint supports (a gather of S*R doubles)."""
    import torch
    import torch.distributed as dist
    from .GPmap import lml_sweep
    rank, world = _rank_world(group)
    S = len(thetas)
    mine = round_robin(S, rank, world)
    local = lml_sweep(X, Y, thetas, indices=mine)
    if world == 1:
        return local
    R = local.shape[1]
    table = torch.zeros((S, R), dtype=torch.float64, device="cuda")
    if mine:
        table[mine] = torch.from_numpy(np.ascontiguousarray(local)).cuda()
    dist.all_reduce(table, group=group)       # disjoint supports: the sum is a gather
    return table.cpu().numpy()
