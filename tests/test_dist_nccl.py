"""NCCL test of the shipped multi-GPU step (bunmpc_b200.dist.ShardedSolver): two ranks on two GPUs, one global batch
sharded interleaved, all_gather of the trajectories and all_reduce of the posterior statistics on the device.  Needs two
visible GPUs (skipped otherwise); the single-GPU variant below checks the same code path with world = 1."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 37            # odd on purpose: the ranks own 19 and 18 instances


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.dist import ShardedSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    batch = synthetic.perturbed(B, seed=4)
    sh = ShardedSolver(batch.n_col, batch.n_eff, shard_batch=(B + world - 1) // world, device=rank)
    sh.upload_global(batch)
    sh.step(params=SolverParams(max_outer=4))
    torch.cuda.synchronize()
    q.put((rank, sh.gathered("F").cpu().numpy()[:B], sh.gathered("X").cpu().numpy()[:B], sh.stats.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _expected(oracle):
    from bunmpc_b200 import synthetic
    from bunmpc_b200.dist import goal_sufficient_stats
    batch = synthetic.perturbed(B, seed=4)
    ref = oracle.solve(batch, params=oracle.default_params(max_outer=4), n_threads=8)
    stats = goal_sufficient_stats(batch.X_ter[:, 3:6], np.nan_to_num(ref["viol"], nan=0.0))
    return ref, stats


def test_sharded_step_two_gpus_nccl(oracle):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, 29700 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, stats = _expected(oracle)
    for rank, F, X, st in res:
        assert np.array_equal(F, ref["F"]) and np.array_equal(X, ref["X"]), rank       # every rank holds the whole batch
        assert st[0] == B and np.allclose(st, stats, rtol=1e-12, atol=1e-15), rank


def test_sharded_step_single_gpu(oracle):
    """world = 1: same step, no collectives; the statistics kernel against numpy."""
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.dist import ShardedSolver
    batch = synthetic.perturbed(B, seed=4)
    sh = ShardedSolver(batch.n_col, batch.n_eff, shard_batch=B, device=0)
    sh.upload_global(batch)
    o = sh.step(params=SolverParams(max_outer=4))
    ref, stats = _expected(oracle)
    assert np.array_equal(o["F"].cpu().numpy(), ref["F"]) and np.array_equal(sh.gathered("X").cpu().numpy(), ref["X"])
    st = sh.stats.cpu().numpy()
    assert st[0] == B and np.allclose(st, stats, rtol=1e-12, atol=1e-15)


# ---- the balanced step: one fresh-instance counter for the whole job (bunmpc_b200.dist.BalancedSolver) ----
BJ = 600          # more instances than resident CTAs per GPU, so instances are parked and resumed as well


def _balanced_worker(rank, world, port, q, exchange="allreduce"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.dist import BalancedSolver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    batch = synthetic.perturbed(BJ, seed=6)
    bs = BalancedSolver(batch.n_col, batch.n_eff, job_batch=BJ, device=rank, exchange=exchange)
    bs.upload_global(batch)
    prm = SolverParams(max_outer=12, slice_outer=3)
    pulled = []
    for _ in range(3):                              # three steps: the two job counters alternate and are cleared in turn
        bs.step(params=prm)
        torch.cuda.synchronize()
        pulled.append(int((bs.dev.out["cycles"] > 0).sum().item()))      # rows this rank solved itself
    q.put((rank, {k: bs.gathered(k).cpu().numpy() for k in ("X", "F", "L", "viol", "iters", "status")},
           bs.stats.cpu().numpy(), pulled))
    dist.barrier()
    bs.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allreduce", "peer"])
def test_balanced_step_two_gpus_nccl(oracle, exchange):
    """Two GPUs pull the instances of one job from one counter (NVLink atomics on CUDA-IPC peer memory); every rank ends
    with every result, bit for bit the oracle's, whichever GPU solved it; together they solved each instance once.
    exchange = "peer": the solve kernel itself stores each finished instance into both GPUs' result rows (fused
    exchange, bunmpc_set_peer_results) instead of an all_reduce of the rows afterwards."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bunmpc_b200 import synthetic
    from bunmpc_b200.dist import goal_sufficient_stats
    world, port = 2, 31700 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port += 7 * (exchange == "peer")
    procs = [ctx.Process(target=_balanced_worker, args=(r, world, port, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = synthetic.perturbed(BJ, seed=6)
    ref = oracle.solve(batch, params=oracle.default_params(max_outer=12), n_threads=8)
    stats = goal_sufficient_stats(batch.X_ter[:, 3:6], np.nan_to_num(ref["viol"], nan=0.0))
    for rank, out, st, pulled in res:
        for k, v in out.items():
            r = ref[k].reshape(v.shape)
            assert np.array_equal(v, r, equal_nan=True), (rank, k)
        assert st[0] == BJ and np.allclose(st, stats, rtol=1e-12, atol=1e-15), rank
    for step in range(3):
        assert sum(p[3][step] for p in res) == BJ, [p[3] for p in res]     # each instance solved exactly once
    print("instances pulled per rank and step:", [p[3] for p in res])   # a rank that starts late may get few


def test_balanced_step_single_gpu(oracle):
    """world = 1: no counter, no exchange; the flat result buffer and its views."""
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.dist import BalancedSolver
    batch = synthetic.perturbed(38, seed=4)
    bs = BalancedSolver(batch.n_col, batch.n_eff, job_batch=38, device=0)
    bs.upload_global(batch)
    o = bs.step(params=SolverParams(max_outer=4))
    ref = oracle.solve(batch, params=oracle.default_params(max_outer=4), n_threads=8)
    for k in ("X", "F", "L", "viol", "iters", "status", "P"):
        assert np.array_equal(o[k].cpu().numpy(), ref[k].reshape(tuple(o[k].shape)), equal_nan=True), k
    bs.close()
