"""world_size-2 gloo test of the multi-process path: interleaved sharding, gather of the solved trajectories
and all-reduce of the posterior sufficient statistics (bunmpc_b200/dist.py).  The per-rank solve is the CPU
oracle here (test stand-in for the local GPU solver; the product path uses BatchSolver.solve)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bunmpc_b200 import dist as bdist, synthetic
    from bunmpc_b200.problem import BatchSolution
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = synthetic.perturbed(5, seed=0)
    prm = oracle.default_params(max_outer=3, max_inner=20)

    def solve_fn(b):
        r = oracle.solve(b, params=prm)
        return BatchSolution(m=b.m, **r)

    sol = bdist.solve_sharded(batch, solve_fn)
    stats = bdist.allreduce_stats(bdist.goal_sufficient_stats(np.full((rank + 1, 3), rank + 1.0), np.ones(rank + 1)))
    # f-3: every rank folds the goals of ITS episodes into the grid posterior; one all-reduce of the log-likelihood grid
    from bunmpc_b200.rollout import GoalPosterior
    post = GoalPosterior(n=12)
    goals = np.array([[0.05, 0.0, 0.0], [0.1, 0.02, -0.02], [0.2, -0.05, 0.03], [0.25, 0.05, 0.0]])
    post.update_batch(goals[bdist.shard_indices(4, rank, world)], all_reduce=True)
    q.put((rank, sol.X, sol.iters, stats, post.p))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_gathers_full_batch():
    sys.path.insert(0, ROOT)
    from bunmpc_b200 import synthetic
    from oracle import oracle
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.solve(synthetic.perturbed(5, seed=0), params=oracle.default_params(max_outer=3, max_inner=20))
    from bunmpc_b200.rollout import GoalPosterior
    one = GoalPosterior(n=12)
    one.update_batch(np.array([[0.05, 0.0, 0.0], [0.1, 0.02, -0.02], [0.2, -0.05, 0.03], [0.25, 0.05, 0.0]]))
    for rank, X, iters, stats, post in res:
        assert np.array_equal(X, full["X"]) and np.array_equal(iters, full["iters"])
        assert stats[0] == 3 and np.allclose(stats[1:4], 1 * 1.0 + 2 * 2.0)
        assert np.allclose(post, one.p, rtol=1e-12)          # sharded posterior == single-process posterior
