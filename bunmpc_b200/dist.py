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


def gather_solutions(local: BatchSolution, B: int, rank: int, world: int, device=None, m=None) -> BatchSolution:
    """all_gather the per-rank results and undo the interleaved sharding: every rank gets the full batch.
    m: the masses of the WHOLE batch ([B] or [1]); without it BatchSolution.mom() of the result raises."""
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
    return BatchSolution(m=m, **out)


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
    return gather_solutions(local, batch.B, rank, world, device=device, m=batch.m)


class ShardedSolver:
    """The multi-GPU step of the path, device resident: ONE global batch, sharded interleaved over the ranks
    (instance i -> rank i % world), solved with no collective on the data path; afterwards the two exchange steps
    BASELINE.json names run over NCCL on the same stream: all_gather of the solved trajectories (F, X) and all_reduce
    of the 17 sufficient statistics of the Bayesian goal update (a fixed-order CUDA reduction of this rank's shard,
    bunmpc_goal_stats_device).  Nothing touches host memory between upload and results.

    One process per GPU (torchrun); with world == 1 the same code runs without the collectives."""

    def __init__(self, n_col: int, n_eff: int = 4, shard_batch: int = 1024, device: int = 0):
        import torch
        import torch.distributed as dist
        from .solver import BatchSolver
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device
        self.solver = BatchSolver(n_col, n_eff, max_batch=shard_batch, device=device)
        self.shard_batch = shard_batch
        dev = torch.device("cuda", device)
        self.stats = torch.zeros(17, dtype=torch.float64, device=dev)
        self._gathered = None
        self.dev = None

    def upload_global(self, batch: CentroidalBatch):
        """Shard the global batch (identical on every rank) and put this rank's shard in HBM."""
        local = batch.shard(self.rank, self.world)
        self.global_B = batch.B
        return self.upload_shard(local)

    def upload_shard(self, local: CentroidalBatch):
        import torch
        if local.B > self.shard_batch:
            raise ValueError("shard larger than the solver was created for")
        self.dev = self.solver.upload(local)
        self.local_B = local.B
        dev = torch.device("cuda", self.device)
        if self.world > 1:
            # every rank holds ceil(B / world) rows in the gather buffers (the last ranks may own one instance less)
            self._gathered = dict(
                F=torch.zeros((self.world * self.shard_batch, self.solver.nf), dtype=torch.float64, device=dev),
                X=torch.zeros((self.world * self.shard_batch, self.solver.nx), dtype=torch.float64, device=dev))
        return self.dev

    def step(self, params=None, arith=0, gather=True, stats=True, goals=None, errors=None):
        """solve + (stats kernel, all_reduce) + (all_gather F, X), all asynchronous on torch's current stream.
        goals: [B_local, 3] device tensor (default: the desired velocity of each instance, X_ter[:, 3:6]);
        errors: [B_local] device tensor (default: the final dynamics violation).  Returns the local output dict."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        out = self.solver.solve_resident(self.dev, params=params, arith=arith)
        if stats:
            g = self.dev.fields["X_ter"] if goals is None else goals
            e = out["viol"] if errors is None else errors
            gin = (_lib.In(g.data_ptr() + 3 * 8, 0 if g.shape[0] == 1 else 9) if goals is None
                   else _lib.In(g.data_ptr(), 3))
            ein = _lib.In(e.data_ptr(), 1)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().bunmpc_goal_stats_device(self.solver._h, int(self.local_B), C.byref(gin), C.byref(ein),
                                                           C.c_void_p(self.stats.data_ptr()), C.c_void_p(stream)),
                       "bunmpc_goal_stats_device")
            if self.world > 1:
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM)
        if gather and self.world > 1:
            nb = self.shard_batch
            for k in ("F", "X"):
                src = out[k] if out[k].shape[0] == nb else torch.nn.functional.pad(out[k], (0, 0, 0, nb - out[k].shape[0]))
                dist.all_gather_into_tensor(self._gathered[k], src)
        return out

    def gathered(self, name: str, ordered: bool = True):
        """The gathered trajectories of the last step() on this rank: [world * shard, width], rank-major; ordered=True
        undoes the interleaved sharding (row i = instance i of the global batch, first global_B rows valid)."""
        if self.world == 1:
            return self.dev.out[name]
        t = self._gathered[name]
        if not ordered:
            return t
        w = t.shape[1]
        return t.view(self.world, self.shard_batch, w).transpose(0, 1).reshape(self.world * self.shard_batch, w)


class BalancedSolver:
    """The multi-GPU step with ONE fresh-instance counter for the whole job (include/bunmpc.h, "multi-GPU jobs").

    A fixed split waits for its slowest shard: the instances' costs spread 4x, so the slowest of G shards of 1024 carries
    about 2 % more work than the average one, and how close a shard gets to its own throughput bound varies by another
    1-2 %.  The instances are independent, so no collective can fix that -- scheduling can: every rank keeps ALL
    instances of the job in its HBM (10.6 KB each), the CTAs of every GPU pull instance ids from a counter in the memory
    of rank 0's GPU (system-scope atomics over NVLink; the counter is mapped into the other ranks' processes by CUDA
    IPC), and a GPU that draws cheap instances simply draws more of them.  Each rank thus fills a SUBSET of the rows of
    its (zeroed) result buffer; one NCCL all_reduce(sum) over the buffers viewed as int64 -- x + 0 is exact for every bit
    pattern, NaN payloads and -0.0 included -- leaves every rank with all results, which is the exchange step
    BASELINE.json names (the all_gather of the fixed split), followed by the same statistics reduction as ShardedSolver.

    Rows are kept rank-major (row r * shard + j  <->  instance j * world + r of the global batch), so a rank's own
    instances are a contiguous slice: `own` below.  One process per GPU; world == 1 degenerates to a plain solve."""

    # result fields that every rank needs after a step, in this order at the front of the flat buffer; P (warm starts)
    # and the cycle counters stay with the rank that solved the row
    EXCHANGED = ("X", "F", "L", "viol", "iters", "status")

    def __init__(self, n_col: int, n_eff: int = 4, job_batch: int = 8192, device: int = 0, exchange: str = "allreduce"):
        """exchange = "allreduce": NCCL all_reduce of the zero-filled result rows after the solve (above);
        exchange = "peer": the fused exchange of include/bunmpc.h (bunmpc_set_peer_results) -- every rank's result rows
        live in a CUDA-IPC buffer that the other ranks map, and the solve kernel's epilogue stores a finished instance
        into the rows of all GPUs; a one-word all_reduce before and after the solve are the only collectives left
        (nobody still reads the previous rows / every rank's kernel has finished)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        from .solver import BatchSolver
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if job_batch % (2 * self.world):
            raise ValueError("job_batch must be a multiple of 2 * world (int32 result fields are exchanged as int64 words)")
        if exchange not in ("allreduce", "peer"):
            raise ValueError("exchange must be 'allreduce' or 'peer'")
        self.exchange = exchange if self.world > 1 else "allreduce"
        self._flat_ptr, self._peer_ptrs = C.c_void_p(), []
        self.device, self.job_batch, self.shard = device, job_batch, job_batch // self.world
        self.solver = BatchSolver(n_col, n_eff, max_batch=job_batch, device=device)
        self.stats = torch.zeros(17, dtype=torch.float64, device=torch.device("cuda", device))
        self._counters = C.c_void_p()
        self.dev = None
        if self.world > 1:
            L, ok, handle = _lib.lib(), 1, None
            if self.rank == 0:
                buf = C.create_string_buffer(64)
                ok = int(L.bunmpc_job_counter_create(device, C.byref(self._counters), buf) == _lib.OK)
                handle = buf.raw
            box = [handle]
            dist.broadcast_object_list(box, src=0)
            if self.rank != 0:
                ok = int(box[0] is not None and
                         L.bunmpc_job_counter_open(device, box[0], C.byref(self._counters)) == _lib.OK)
            flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", device))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) != 1:
                self.close()
                raise RuntimeError("BalancedSolver: the job counter could not be mapped into every rank's process "
                                   "(CUDA IPC): " + (L.bunmpc_last_error() or b"").decode())
            _lib.check(L.bunmpc_set_job_counter(self.solver._h, self._counters, int(self.rank == 0)), "bunmpc_set_job_counter")

    def close(self):
        from . import _lib
        if getattr(self, "_peer_ptrs", None) or (getattr(self, "_flat_ptr", None) is not None and self._flat_ptr.value):
            import torch
            torch.cuda.synchronize(self.device)
            if getattr(self, "solver", None) is not None and self.solver._h.value:
                _lib.lib().bunmpc_set_peer_results(self.solver._h, 0, None)
            for p in self._peer_ptrs:
                _lib.lib().bunmpc_peer_buffer_release(p, 0)
            self._peer_ptrs = []
            self.flat = None
            if self.dev is not None:
                self.dev.out = {}
            if self._flat_ptr.value:
                try:        # the peers close their mappings of this buffer before it is freed (close() is collective)
                    import torch.distributed as dist
                    if dist.is_initialized() and self.world > 1:
                        dist.barrier()
                except Exception:
                    pass
                _lib.lib().bunmpc_peer_buffer_release(self._flat_ptr, 1)
                self._flat_ptr.value = None
        if getattr(self, "_counters", None) is not None and self._counters.value:
            if getattr(self, "solver", None) is not None and self.solver._h.value:
                _lib.lib().bunmpc_set_job_counter(self.solver._h, None, 0)
            _lib.lib().bunmpc_job_counter_release(self._counters, int(self.rank == 0))
            self._counters.value = None
        if getattr(self, "solver", None) is not None:
            self.solver.close()

    def row_order(self):
        """row -> instance of the global batch (rank-major rows of the interleaved split)"""
        return np.concatenate([np.arange(r, self.job_batch, self.world) for r in range(self.world)])

    @property
    def own(self):
        return slice(self.rank * self.shard, (self.rank + 1) * self.shard)

    def upload_global(self, batch: CentroidalBatch):
        """Every rank puts the WHOLE job in its HBM (identical `batch` on every rank)."""
        import torch
        if batch.B != self.job_batch:
            raise ValueError("batch size != job_batch")
        self.dev = self.solver.upload(batch.select(self.row_order()))
        B, dev = self.job_batch, torch.device("cuda", self.device)
        shapes = dict(X=((B, self.solver.nx), torch.float64), F=((B, self.solver.nf), torch.float64), L=((B, 2), torch.float64),
                      viol=((B,), torch.float64), iters=((B, 5), torch.int32), status=((B,), torch.int32),
                      P=((B, self.solver.nx), torch.float64), cycles=((B,), torch.int64))
        words = {k: int(np.prod(sh)) * (1 if dt != torch.int32 else 0) + (int(np.prod(sh)) // 2 if dt == torch.int32 else 0)
                 for k, (sh, dt) in shapes.items()}
        n_words = sum(words.values())
        if self.exchange == "peer":
            self.flat = self._peer_flat(n_words)
        else:
            self.flat = torch.zeros(n_words, dtype=torch.int64, device=dev)
        off, offs = 0, {}
        for k in self.EXCHANGED + ("P", "cycles"):
            sh, dt = shapes[k]
            self.dev.out[k] = self.flat[off:off + words[k]].view(dt).view(sh)
            offs[k] = 8 * off
            off += words[k]
            if k == self.EXCHANGED[-1]:
                self.n_exchanged = off
        if self.exchange == "peer":
            self._register_peers(offs)
            self._word = torch.zeros(1, dtype=torch.int32, device=dev)
        return self.dev

    def _peer_flat(self, n_words: int):
        """The flat result buffer as a CUDA-IPC allocation of the library, wrapped as a torch tensor; the other ranks'
        buffers mapped into this process."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        L, buf = _lib.lib(), C.create_string_buffer(64)
        ok = int(L.bunmpc_peer_buffer_create(self.device, 8 * n_words, C.byref(self._flat_ptr), buf) == _lib.OK)
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.raw if ok else None)
        if ok and all(h is not None for h in handles):
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                if L.bunmpc_peer_buffer_open(self.device, h, C.byref(p)) != _lib.OK:
                    ok = 0
                    break
                self._peer_ptrs.append(p)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            err = (L.bunmpc_last_error() or b"").decode()
            self.close()
            raise RuntimeError("BalancedSolver: the result rows could not be mapped between the ranks (CUDA IPC): " + err)

        class _Raw:       # torch.as_tensor reads __cuda_array_interface__
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (int(self._flat_ptr.value), False),
                                        "version": 2}
        self._raw = raw
        return torch.as_tensor(raw, device=torch.device("cuda", self.device))

    def _register_peers(self, offs: dict):
        import ctypes as C

        from . import _lib
        arr = (_lib.Solution * len(self._peer_ptrs))()
        for g, p in enumerate(self._peer_ptrs):
            base = int(p.value)
            arr[g] = _lib.Solution(base + offs["X"], base + offs["F"], None, base + offs["L"], base + offs["iters"],
                                   base + offs["viol"], base + offs["status"], None, None)
        _lib.check(_lib.lib().bunmpc_set_peer_results(self.solver._h, len(self._peer_ptrs), arr), "bunmpc_set_peer_results")

    def load_own_rows(self, pinned: dict):
        """e2e: this rank's instances arrive from ITS host (pinned tensors of its shard, [shard, w] per per-instance
        field), go up into its rows and reach the other ranks over NVLink (one in-place all_gather per field)."""
        import torch.distributed as dist
        for f, t in pinned.items():
            g = self.dev.fields[f]
            if g is None or g.shape[0] == 1:
                continue
            g[self.own].copy_(t, non_blocking=True)
            if self.world > 1:
                dist.all_gather_into_tensor(g, g[self.own])

    def step(self, params=None, arith=0, stats=True):
        """zero the result rows, solve what this GPU pulls, exchange; returns the dict of FULL result tensors
        (X, F, L, viol, iters, status: all rows on every rank; P, cycles: the rows this rank solved)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        if self.exchange == "peer":
            dist.all_reduce(self._word)                    # no rank still reads the rows of the previous step
            self.flat[self.n_exchanged:].zero_()           # P and the cycle counters: set on the rank that solves the row
            out = self.solver.solve_resident(self.dev, params=params, arith=arith)   # stores every row on every GPU
            dist.all_reduce(self._word)                    # every rank's kernel has finished: all rows are here
        else:
            self.flat.zero_()
            out = self.solver.solve_resident(self.dev, params=params, arith=arith)
            if self.world > 1:
                dist.all_reduce(self.flat[:self.n_exchanged], op=dist.ReduceOp.SUM)
        if stats:
            g, e = self.dev.fields["X_ter"], out["viol"]
            o = self.rank * self.shard
            gin = _lib.In(g.data_ptr() + 3 * 8 + (0 if g.shape[0] == 1 else 9 * 8 * o), 0 if g.shape[0] == 1 else 9)
            ein = _lib.In(e.data_ptr() + 8 * o, 1)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().bunmpc_goal_stats_device(self.solver._h, int(self.shard), C.byref(gin), C.byref(ein),
                                                           C.c_void_p(self.stats.data_ptr()), C.c_void_p(stream)),
                       "bunmpc_goal_stats_device")
            if self.world > 1:
                dist.all_reduce(self.stats, op=dist.ReduceOp.SUM)
        return out

    def gathered(self, name: str, ordered: bool = True):
        """Results of the last step on this rank, all rows; ordered=True: row i = instance i of the global batch."""
        t = self.dev.out[name]
        if not ordered or self.world == 1:
            return t
        w = t.shape[1] if t.dim() > 1 else 1
        return t.view(self.world, self.shard, w).transpose(0, 1).reshape(self.job_batch, w)


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
