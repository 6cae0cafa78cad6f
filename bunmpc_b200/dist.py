"""Multi-GPU use of the batched solver: one process per GPU, instances sharded, results gathered.

The path shards trivially -- MPC instances never exchange data during a solve (the reference runs them
strictly one after the other, iterative_algorithm/data_collection.py:181-277) -- so there is NO collective on
the data path.  The only exchange steps are the ones BASELINE.json names: gathering the solved trajectories
and reducing the sufficient statistics of the Bayesian goal update (locosafedagger_modified.py:357-402),
both over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Sharding is interleaved (instance i -> rank i % world) so that the 4x spread in iteration counts between
instances is spread over the ranks as well.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from .problem import BatchSolution, CentroidalBatch


def shard_indices(B: int, rank: int, world: int) -> np.ndarray:
    return np.arange(rank, B, world)


def _dist():
    import torch.distributed as dist
    return dist


def gather_solutions(local: BatchSolution, B: int, rank: int, world: int, device=None) -> BatchSolution:
    """all_gather the per-rank results and undo the interleaved sharding: every rank gets the full batch."""
    import torch
    dist = _dist()
    per = (B + world - 1) // world                       # padded shard size
    out = {}
    for name in ("X", "F", "P", "L", "iters", "viol", "status"):
        a = getattr(local, name)
        a2 = a.reshape(a.shape[0], -1)
        pad = np.zeros((per, a2.shape[1]), dtype=a2.dtype)
        pad[: a2.shape[0]] = a2
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        full = torch.empty((world * per, a2.shape[1]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t)
        full = full.cpu().numpy().reshape(world, per, -1)
        res = np.empty((B, a2.shape[1]), dtype=a2.dtype)
        for r in range(world):
            idx = shard_indices(B, r, world)
            res[idx] = full[r, : len(idx)]
        out[name] = res.reshape((B,) + a.shape[1:])
    return BatchSolution(m=None, **out)


def solve_sharded(batch: CentroidalBatch, solve_fn: Callable[[CentroidalBatch], BatchSolution],
                  gather: bool = True, device=None) -> BatchSolution:
    """Solve this rank's interleaved shard of `batch` with solve_fn (normally BatchSolver.solve on the local GPU)
    and, if gather, return the whole batch's solution on every rank."""
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local = solve_fn(batch.shard(rank, world))
    if not gather or world == 1:
        return local
    return gather_solutions(local, batch.B, rank, world, device=device)


# ---------------------------------------------------------------------------------------------------
# Bayesian goal update (grid posterior over velocity goals), sufficient statistics over ranks
# ---------------------------------------------------------------------------------------------------
def goal_sufficient_stats(goals: np.ndarray, errors: np.ndarray) -> np.ndarray:
    """Per-shard sufficient statistics of (goal g_i in R^3, error e_i): [N, sum g (3), sum g g^T (9), sum e,
    sum e g (3)] -> 17 doubles, summed over ranks by allreduce_stats()."""
    g = np.asarray(goals, dtype=np.float64).reshape(-1, 3)
    e = np.asarray(errors, dtype=np.float64).reshape(-1)
    return np.concatenate([[g.shape[0]], g.sum(0), (g[:, :, None] * g[:, None, :]).sum(0).ravel(), [e.sum()],
                           (e[:, None] * g).sum(0)])


def allreduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(stats, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gaussian_likelihood_grid(grid_axes, observed_goal, sigma=0.1):
    """Likelihood of locosafedagger_modified.py:357-384: independent Gaussians centred at the observed goal."""
    vx, vy, w = np.meshgrid(*grid_axes, indexing="ij")
    g = np.asarray(observed_goal, dtype=np.float64)
    return np.exp(-0.5 * (((vx - g[0]) / sigma) ** 2 + ((vy - g[1]) / sigma) ** 2 + ((w - g[2]) / sigma) ** 2))


def posterior_update(prior: np.ndarray, likelihood: np.ndarray) -> np.ndarray:
    """posterior ~ prior * likelihood, normalised (locosafedagger_modified.py:386-402)."""
    post = prior * likelihood
    s = post.sum()
    return post / s if s > 0 else np.full_like(post, 1.0 / post.size)
