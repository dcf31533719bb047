"""Multi-GPU sharding: one process per GPU (torchrun), torch.distributed for the plumbing.

The path shards without any data-path collective: independent paths, grid rows
and hyper-parameter points are split across ranks; the single-matrix Cholesky is replicated on every
rank (deterministic, so every rank holds the same factor -- no broadcast needed).  The only collective
is the final all-gather of the results (NCCL on GPUs; gloo in the CPU tests of the index logic).
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


def all_gather_rows(local, counts, group=None):
    """All-gather tensors that differ only in their first dimension (counts[r] rows on rank r).

    Returns the concatenation in rank order on every rank.  Works for CUDA tensors over NCCL and CPU
    tensors over gloo.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    tail = tuple(local.shape[1:])
    cmax = max(counts)
    if len(set(counts)) == 1:
        flat = torch.empty((sum(counts),) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
        return flat
    # uneven shares: pad every rank's block to the largest count (gloo and NCCL both want equal sizes)
    padded = torch.zeros((cmax,) + tail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    flat = torch.empty((world * cmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, padded, group=group)
    return torch.cat([flat[r * cmax: r * cmax + c] for r, c in enumerate(counts)], dim=0)


def predict_grid_sharded(model, bounds, shape, t=None, return_var=True, include_noise=False, group=None,
                         gather=True):
    """Each rank evaluates its contiguous share of the Gx*Gy grid points; results are all-gathered."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    Gx, Gy = int(shape[0]), int(shape[1])
    lo, hi = shard_range(Gx * Gy, rank, world)
    out = model.predict_grid(bounds, shape, t=t, return_var=return_var, include_noise=include_noise, points=(lo, hi))
    if not gather or world == 1:
        return out
    counts = shard_counts(Gx * Gy, world)
    if return_var:
        mu = all_gather_rows(out[0], counts, group).view(Gy, Gx, -1)
        var = all_gather_rows(out[1], counts, group).view(Gy, Gx)
        return mu, var
    return all_gather_rows(out, counts, group).view(Gy, Gx, -1)


def fit_gp_batched_sharded(Xb_local, Yb_local, counts, group=None, gather=True, **kw):
    """Each rank fits its own paths (already local); alpha / lml are all-gathered in rank order."""
    from .GPmap import fit_gp_batched
    alpha, lml = fit_gp_batched(Xb_local, Yb_local, **kw)
    if not gather:
        return alpha, lml
    return all_gather_rows(alpha, counts, group), all_gather_rows(lml, counts, group)


def lml_sweep_sharded(X, Y, thetas, group=None):
    """Round-robin the hyper-parameter points over ranks, all-gather the S x R table."""
    import torch
    import torch.distributed as dist
    from .GPmap import lml_sweep
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
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
