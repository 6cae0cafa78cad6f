#!/usr/bin/env python
"""Benchmark of the centroidal MPC hot path (BASELINE.json metric: centroidal MPC solves/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (create_* + BiConvexMP::optimize for every instance) over one batch of
synthetic perturbed Solo12-trot states: BASELINE config[1], 1024 instances per GPU, horizon 20.  With N GPUs the
step is bunmpc_b200.dist.ShardedSolver.step on ONE global batch of 1024 N instances sharded interleaved (instance
i -> rank i mod N): solve, then the path's two exchange steps over NCCL -- all_gather of the solved trajectories and
all_reduce of the 17 posterior sufficient statistics.
  value : solves/s with the batch resident in HBM (CUDA events on the launching stream, L2 flushed
          between steps, max over ranks)
  e2e   : the same through the host-buffer C ABI call (pinned host inputs -> H2D -> kernels -> D2H)
  roofline : algorithmic FP64 work of the solve kernel / its event time, against the DFMA peak measured
          in this run (the path is FP64-pipe/latency bound, not HBM bound; the HBM view is added as roofline_hbm)
  cpu_baseline : the CPU oracle (a port of the reference algorithm) on this box's host cores, same inputs
--impl reference times that CPU implementation alone (the reference itself cannot be built here: no Eigen).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "centroidal_mpc_solves_per_sec"
UNIT = "solves/s"
B_PER_GPU = 1024
WORKLOAD = "solo12_trot_perturbed_B1024_n20 (BASELINE config[1])"


def algorithmic_flops(n_col, iters):
    """FLOPs(solve) = n (1870 K_out + 730 I_F + 471 I_X), mul and add counted separately."""
    it = np.asarray(iters, dtype=np.float64)
    return float(n_col * (1870.0 * it[:, 0].sum() + 730.0 * it[:, 1].sum() + 471.0 * it[:, 2].sum()))


def algorithmic_bytes(n_col, B):
    """compulsory HBM traffic per solve = (62 n + 80) * 8 bytes."""
    return float(B) * (62 * n_col + 80) * 8


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu(batch, cores, max_instances, budget_s=25.0):
    """Oracle on the host cores over a bounded sample of the batch; returns solves/s and what was run."""
    from oracle import oracle
    # size the sample from a short probe so the leg stays within the budget
    probe = batch.select(np.arange(min(cores, batch.B)))
    t0 = time.perf_counter()
    oracle.solve(probe, n_threads=cores)
    per_wave = max(time.perf_counter() - t0, 1e-3)
    n = int(min(max_instances, batch.B, max(cores, cores * int(budget_s / per_wave))))
    sample = batch.select(np.arange(n))
    ex = oracle.expand(sample)                       # builders outside the timed region, like the GPU "value"
    x_init = np.broadcast_to(sample.x_init, (n, 9))
    nx, nf = sample.nx, sample.nf
    X0, F0, P0 = np.tile(x_init, (1, sample.n_col + 1)), np.zeros((n, nf)), np.zeros((n, nx))
    t0 = time.perf_counter()
    oracle.solve_expanded(sample.n_col, sample.n_eff, sample.m, sample.rho, x_init, sample.cnt_plan, sample.dt,
                          ex["Qx"], ex["qx"], ex["Qf"], ex["qf"], ex["lbx"], ex["ubx"], X0, F0, P0, sample.L0,
                          n_threads=cores)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def ncu_constants():
    """Per-launch counters of the profiled solve kernel, written by profiles/summarize.py from an ncu report of THIS
    workload (with the commit of the kernel they belong to); None when the file is missing."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_constants.json")))
    except Exception:
        return None


def smem_roofline(iters, k_time, clocks, info, ncu):
    """Shared-memory data pipe: wavefronts per inner iteration from the ncu capture (l1tex__data_pipe_lsu_wavefronts_
    mem_shared of the profiled launch / its inner iterations), scaled by the live iteration counters and kernel time."""
    mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
    if not ncu or not mhz or not k_time or not ncu.get("smem_wavefronts_per_inner_iteration"):
        return None
    wf_per_iter = float(ncu["smem_wavefronts_per_inner_iteration"])
    wf = wf_per_iter * float(iters[:, 1].sum() + iters[:, 2].sum())
    ach = wf / (k_time * mhz * 1e6 * info["num_sms"])
    return {"bound": "shared-memory data pipe", "achieved": ach, "peak": 1.0, "unit": "wavefronts/clk/SM", "frac": ach,
            "peak_source": "nominal 1 wavefront (128 B) per clock per SM",
            "wavefronts_per_inner_iteration": wf_per_iter, "source": ncu.get("source")}


def other_workloads(device, arith, world=1, rank=0):
    """Short runs of the other BASELINE.json configs at their real sizes: total batch / world instances per GPU
    (16 384 Go2 trot, 4 096 bound n=48 / jump n=60, 65 536 Bayes samples over 8 GPUs -> 2048 / 512 / 8192 per GPU; a
    single-GPU run measures that per-GPU shard).  Device-resident ShardedSolver.step (solve + statistics + gathers under
    world > 1), one warm-up + one timed step each, max over ranks: reported for context, not part of `value`."""
    import torch
    import torch.distributed as dist
    from bunmpc_b200 import SolverParams, synthetic
    from bunmpc_b200.dist import ShardedSolver
    from bunmpc_b200.motions import GO2_RETUNE_SOLVER
    out = {}
    seed = 100 + rank     # the global batch of a config = the union of the per-rank shards (i.i.d. instances)
    retuned = SolverParams(**GO2_RETUNE_SOLVER)
    cases = [("config2_go2_trot_16384 (Solo12 gait records, Go2 mass: the reference algorithm diverges, DESIGN.md)", 2048, lambda B: synthetic.config(2, B=B, seed=seed), None),
             ("config2_go2_trot_16384_RETUNED (non-reference weights, mu, exit_tol: motions.GO2_RETUNE_*)", 2048, lambda B: synthetic.perturbed(B, "go2_retuned", "trot", seed=seed), retuned),
             ("config3_go2_bound_n48_4096 (reference algorithm diverges)", 512, lambda B: synthetic.config(3, B=B, seed=seed), None),
             ("config3_go2_bound_n48_4096_RETUNED (non-reference)", 512, lambda B: synthetic.perturbed(B, "go2_retuned", "bound", seed=seed, horizon_scale=2.0), retuned),
             ("config3_solo12_bound_n48_4096 (config 3 gait and horizon, Solo12 mass)", 512, lambda B: synthetic.perturbed(B, "solo12", "bound", seed=seed, horizon_scale=2.0), None),
             ("config3_solo12_jump_n60_4096", 512, lambda B: synthetic.perturbed(B, "solo12", "jump", seed=seed, horizon_scale=2.0), None),
             ("solo12_bound_n24_8192 (bound gait, its own horizon)", 1024, lambda B: synthetic.perturbed(B, "solo12", "bound", seed=seed), None),
             ("solo12_jump_n30_8192 (jump gait, its own horizon)", 1024, lambda B: synthetic.perturbed(B, "solo12", "jump", seed=seed), None),
             ("acyclic_jump_fwd_n25_8192 (SoloAcyclicGen replans over the whole motion, 50 outer iterations)", 1024, lambda B: synthetic.acyclic_replans(B, "jump_fwd", seed=seed), SolverParams(max_outer=50)),
             ("acyclic_rearing_n20_8192 (50 outer iterations; replans from a standing state into the rearing phase end in NaN in the reference algorithm too)", 1024, lambda B: synthetic.acyclic_replans(B, "rearing", seed=seed), SolverParams(max_outer=50)),
             ("config4_bayes_goal+weight_samples_65536", 8192, lambda B: synthetic.config(4, B=B, seed=seed), None),
             ("solo12_trot_65536 (saturated)", 8192, lambda B: synthetic.config(1, B=B, seed=seed), None)]
    for name, per_gpu, make, prm in cases:
        b = make(per_gpu)
        sh = ShardedSolver(b.n_col, b.n_eff, shard_batch=per_gpu, device=device)
        sh.upload_shard(b)
        sh.step(params=prm, arith=arith)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = sh.step(params=prm, arith=arith)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
        st = o["status"]
        agg = torch.tensor([float((st == 0).sum()), float((st == 2).sum()), float(o["iters"][:, 0].sum()),
                            float(o["iters"][:, 1:3].sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        tot = world * per_gpu
        agg = agg.cpu().numpy()
        out[name] = {"solves_per_s": tot / float(tt.item()), "ms": 1e3 * float(tt.item()), "instances": tot,
                     "per_gpu": per_gpu, "n_col": b.n_col, "outer_mean": agg[2] / tot, "inner_mean": agg[3] / tot,
                     "converged_frac": agg[0] / tot, "nan_frac": agg[1] / tot,
                     "posterior_stats_N": float(sh.stats[0].item())}
        sh.solver.close()
    if world > 1:
        return out
    # f-1: only the centroidal states cross PCIe; contact plan and references are built on the device
    from bunmpc_b200.motions import GAITS, ROBOTS
    from bunmpc_b200.solver import BatchSolver
    rb, gp = ROBOTS["solo12"], GAITS["solo12"]["trot"]
    rng = np.random.default_rng(0)
    B = 1024
    com = np.array([0.0, 0.0, gp.nom_ht]) + rng.normal(0.0, 0.02, (B, 3))
    vcom, amom = rng.normal(0.0, 0.1, (B, 3)), rng.normal(0.0, 0.02, (B, 3))
    foot = np.broadcast_to(rb.foot_pos, (B, 4, 3)) + np.concatenate([rng.normal(0.0, 0.02, (B, 4, 2)), np.zeros((B, 4, 1))], 2)
    t0s = rng.integers(0, 10, B) * gp.gait_dt
    v_des = np.stack([rng.uniform(0.0, 0.3, B), np.zeros(B), np.zeros(B)], 1)
    s = BatchSolver(gp.horizon(), 4, max_batch=B, device=device)

    def run():
        dev = s.build_device(rb, gp, com, vcom, amom, foot, t0s, v_des, np.zeros(B))
        o = s.solve_resident(dev, arith=arith)
        return o["F"].cpu(), o["X"].cpu(), o["status"].cpu()
    run()
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    out["solo12_trot_B1024_from_centroidal_states (device-side problem builder, e2e)"] = {
        "solves_per_s": B / dt, "ms": 1e3 * dt, "h2d_bytes": int(B * 8 * (9 + 12 + 1 + 3 + 1 + 2))}
    s.close()
    # f-3: lock-step rollouts, one batched solve per replanning tick, step sizes carried, host-side plan builder and plant
    from bunmpc_b200.rollout import EpisodeState, LockstepRollouts, TrackingPlant
    st = EpisodeState(com, vcom, amom, np.array(foot), t0s.astype(np.float64), np.zeros(B))
    roll = LockstepRollouts(rb, gp, plant=TrackingPlant(0.002, 0.02, 0.0, seed=1), device=device)
    roll.run(st, v_des, 0.0, n_ticks=1)
    roll.h2d_bytes = 0
    t0 = time.perf_counter()
    rec = roll.run(st, v_des, 0.0, n_ticks=6)
    dt = time.perf_counter() - t0
    out["lockstep_rollouts_B1024_6_ticks (host plan builder + plant in the loop, e2e)"] = {
        "solves_per_s": int((np.array([(r > 0).any(1).sum() for r in rec.iters])).sum()) / dt, "ms": 1e3 * dt,
        "failed": int((rec.failed_at >= 0).sum()), "h2d_bytes": int(roll.h2d_bytes)}
    # the same loop on the device path: states up, build_problem_kernel + solve where the problem lies, plan down
    rold = LockstepRollouts(rb, gp, plant=TrackingPlant(0.002, 0.02, 0.0, seed=1), device=device, builder="device")
    rold.run(st, v_des, 0.0, n_ticks=1)
    rold.h2d_bytes = 0
    t0 = time.perf_counter()
    rec = rold.run(st, v_des, 0.0, n_ticks=6)
    dt = time.perf_counter() - t0
    out["lockstep_rollouts_B1024_6_ticks_device_builder (states up, plan down, plant in the loop, e2e)"] = {
        "solves_per_s": int((np.array([(r > 0).any(1).sum() for r in rec.iters])).sum()) / dt, "ms": 1e3 * dt,
        "failed": int((rec.failed_at >= 0).sum()), "h2d_bytes": int(rold.h2d_bytes)}
    return out


def cpu_single_solve_latency(batch, n=96):
    """p50 / p95 of ONE solve on ONE host core (the reference's own usage: one BiConvexMP::optimize per replan,
    comparable to its dyn_time stamp, kino_dyn.cpp:46-48), over the first n instances of the workload."""
    from oracle import oracle
    lat = []
    for i in range(min(n, batch.B)):
        one = batch.select(np.arange(i, i + 1))
        t0 = time.perf_counter()
        oracle.solve(one, n_threads=1)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e3
    return {"p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)), "solves": int(len(lat)),
            "cores": 1}


def reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (oracle port; the reference's own
    sources need Eigen, which this image does not have).  Rank 0 alone runs and prints."""
    if rank != 0:
        return
    from bunmpc_b200 import synthetic
    cores = cpu_cores()
    batch = synthetic.config(1, B=B_PER_GPU, seed=0)
    per_step = min(batch.B, 16 * cores)
    for _ in range(args.warmup):
        run_cpu(batch.select(np.arange(cores)), cores, cores, budget_s=1.0)
    done, t_total = 0, 0.0
    for k in range(args.steps):
        idx = (np.arange(per_step) + k * per_step) % batch.B
        v, n, dt = run_cpu(batch.select(idx), cores, per_step, budget_s=60.0)
        done += n; t_total += dt
    value = done / t_total
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_step": per_step, "n_col": batch.n_col},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} instances of the workload per step, {cores} threads, "
                                       "oracle/bicon_oracle.c (gcc -O3), solve only"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="instances per GPU per step")
    ap.add_argument("--arith", default="strict", choices=["strict", "fma", "mixed"])
    ap.add_argument("--split", default="fixed", choices=["fixed", "balanced", "fused"],
                    help="multi-GPU only: 'fixed' = instance i -> rank i mod G (ShardedSolver, the default: measured equal or "
                         "faster at 2 and 8 GPUs, DESIGN.md section 7), 'balanced' = one fresh-instance counter for all GPUs "
                         "over NVLink peer memory (bunmpc_b200.dist.BalancedSolver), 'fused' = the same counter and the solve "
                         "kernel stores every finished instance into the result rows of ALL GPUs (no collective on the results)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from bunmpc_b200 import synthetic, ARITH_FMA, ARITH_MIXED, ARITH_STRICT
    from bunmpc_b200.dist import BalancedSolver, ShardedSolver

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    arith = {"strict": ARITH_STRICT, "fma": ARITH_FMA, "mixed": ARITH_MIXED}[args.arith]
    B = args.batch

    # ONE global batch of world * B perturbed states (the same on every rank), sharded interleaved: instance i -> rank
    # i mod world.  Weak scaling: B instances per GPU at every N.
    global_batch = synthetic.config(1, B=world * B, seed=0)
    balanced = None
    if world > 1 and args.split in ("balanced", "fused"):
        try:
            balanced = BalancedSolver(global_batch.n_col, global_batch.n_eff, job_batch=world * B, device=local_rank,
                                      exchange="peer" if args.split == "fused" else "allreduce")
        except RuntimeError as exc:        # every rank raises or none does (the ranks agree inside the constructor)
            if rank == 0:
                print(f"bench: balanced split unavailable ({exc}); using the fixed split", file=sys.stderr)
    if balanced is not None:
        dev = balanced.upload_global(global_batch)          # the whole job on every GPU, rank-major rows
        sharded, solver = None, balanced.solver
    else:
        sharded = ShardedSolver(global_batch.n_col, global_batch.n_eff, shard_batch=B, device=local_rank)
        dev = sharded.upload_global(global_batch)
        solver = sharded.solver
    batch = global_batch.shard(rank, world)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_resident():
        # the package's multi-GPU step.  fixed split: solve, statistics kernel + all_reduce, all_gather of F and X;
        # balanced: zero the result rows, solve what this GPU pulls from the job counter, all_reduce of the rows, statistics
        return balanced.step(arith=arith) if balanced is not None else sharded.step(arith=arith)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()

    # ---- timed: K steps, device time per step (events on the launching stream), L2 flushed in between ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = solver.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        step_resident()
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = solver.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    t_dev = float(step_ms.sum()) * 1e-3
    tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    value = world * B * args.steps / t_max

    out = {k: v.cpu().numpy() for k, v in dev.out.items()}
    if balanced is not None:        # the rows this GPU solved in the last step (its cycle counters are set there only)
        took = out["cycles"] > 0
        out = {k: v[took] for k, v in out.items()}
    iters = out["iters"]

    # ---- solve-kernel time alone (events around the solve launch only), for the roofline ----
    kt = []
    kflops = []
    for _ in range(max(3, min(args.steps, 5))):
        flush.zero_()
        if balanced is not None:
            balanced.flat.zero_()
            barrier()               # two solves of a job are separated by a collective (the job counters alternate)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        solver.solve_resident(dev, arith=arith)
        b.record(stream)
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b) * 1e-3)
        if balanced is not None:    # which instances a GPU pulls differs from launch to launch: count them per launch
            tk = dev.out["cycles"] > 0
            kflops.append(algorithmic_flops(batch.n_col, dev.out["iters"][tk].cpu().numpy()))
    k_time = float(np.mean(kt))
    # the same on every rank: how unequal the shards are (a step ends with collectives, i.e. with the slowest rank), and
    # how much work each shard holds
    rank_info = None
    if world > 1:
        mine = torch.tensor([1e3 * k_time, float(iters[:, 1:3].sum())], dtype=torch.float64, device="cuda")
        allr = torch.zeros((world, 2), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().numpy()
        rank_info = {"solve_kernel_ms_per_rank": [round(float(v), 3) for v in allr[:, 0]],
                     "inner_iterations_per_rank": [int(v) for v in allr[:, 1]],
                     "note": "a step takes the slowest rank's solve plus the exchange; mean / max of the kernel times is "
                             "the ceiling of the weak-scaling efficiency for i.i.d. shards"}

    # ---- e2e: host buffers in, host buffers out, every copy inside the timed region ----
    # N = 1: the reference-facing C ABI call bunmpc_solve_compact_host (pinned inputs -> H2D -> kernels -> D2H).
    # N > 1: each rank copies its pinned shard into the device batch, runs ShardedSolver.step (solve + statistics +
    # NCCL gathers) and reads its results and the reduced statistics back.
    import ctypes as C
    from bunmpc_b200 import _lib
    pinned = {}
    for f in _lib.COMPACT_FIELDS:
        a = getattr(batch, f)
        if a is None:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a.reshape(a.shape[0], -1))).pin_memory()
        pinned[f] = t
        setattr(batch, f, t.numpy().reshape(a.shape))
    if balanced is not None:        # every rank reads back the rows of ITS instances (all rows are on every GPU after a step)
        outbuf = {k: torch.empty(dev.out[k][balanced.own].shape, dtype=dev.out[k].dtype).pin_memory()
                  for k in BalancedSolver.EXCHANGED}
    else:
        outbuf = {k: torch.empty(v.shape, dtype=torch.from_numpy(v).dtype).pin_memory() for k, v in out.items()}
    outnp = {k: v.numpy() for k, v in outbuf.items()}
    stats_host = torch.empty(17, dtype=torch.float64).pin_memory()

    def e2e_step():
        if world == 1:
            solver.solve(batch, arith=arith, out=outnp)
            return
        if balanced is not None:    # own instances up, to the other GPUs over NVLink, step, own rows down
            balanced.load_own_rows(pinned)
            o = balanced.step(arith=arith)
            for k, v in outbuf.items():
                v.copy_(o[k][balanced.own], non_blocking=True)
            stats_host.copy_(balanced.stats, non_blocking=True)
            torch.cuda.synchronize()
            return
        for f, t in pinned.items():
            dev.fields[f].copy_(t, non_blocking=True)
        o = sharded.step(arith=arith)
        for k, v in outbuf.items():
            v.copy_(o[k], non_blocking=True)
        stats_host.copy_(sharded.stats, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    h2d = batch.input_bytes()
    d2h = int(sum(v.numel() * v.element_size() for v in outbuf.values())) + (17 * 8 if world > 1 else 0)

    if balanced is not None:        # what follows are this rank's own solves: back to the local instance counter
        barrier()
        _lib.check(_lib.lib().bunmpc_set_job_counter(solver._h, None, 0), "bunmpc_set_job_counter")

    # ---- p50 latency of a single solve through the host API (B = 1) ----
    one = batch.select(np.arange(1))
    lat = []
    for i in range(25):
        t0 = time.perf_counter()
        solver.solve(one, arith=arith)
        lat.append(time.perf_counter() - t0)
    p50_ms = float(np.median(lat[5:]) * 1e3)

    # the other BASELINE configs at their real sizes (every rank takes part when world > 1)
    others = None
    if not args.no_extra:
        others = other_workloads(local_rank, arith, world, rank)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the solve kernel ----
    fp64_peak = solver.measure_fp64_peak()
    flops = float(np.mean(kflops)) if kflops else algorithmic_flops(batch.n_col, iters)
    achieved = flops / k_time * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = algorithmic_bytes(batch.n_col, B) / k_time * 1e-9
    info = solver.kernel_info()
    ncu = ncu_constants()
    traffic = ncu.get("dram_bytes_per_launch") if (ncu and B == 1024) else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.arith != "mixed" else "f64 arithmetic, f32 storage of the Hessian / constraint rows",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_gpu_per_step": B, "n_col": batch.n_col, "n_eff": batch.n_eff,
                   "arith": args.arith, "l2": "flushed between steps (256 MiB write)",
                   "sharding": (f"one global batch of {world * B} instances resident on every GPU; the CTAs of all GPUs pull "
                                "instance ids from ONE counter in rank 0's HBM (system-scope atomics over NVLink peer "
                                "memory, CUDA IPC); per step: " + ("the solve kernel stores each finished instance into the "
                                "result rows of every GPU (peer stores over NVLink), two one-word all_reduce as barriers, "
                                if args.split == "fused" else "NCCL all_reduce of the zero-filled result rows (exact), ") +
                                "all_reduce of 17 posterior statistics (bunmpc_b200.dist.BalancedSolver)"
                                if balanced is not None else
                                f"one global batch of {world * B} instances, instance i -> rank i mod {world}" + ("; per step: NCCL all_gather of F and X, all_reduce of 17 posterior statistics (bunmpc_b200.dist.ShardedSolver)" if world > 1 else "")),
                   **({"split": args.split if balanced is not None else "fixed"} if world > 1 else {}),
                   "kernel": info},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "p50_latency_ms": p50_ms,
        "wall_s_timed_region": t_wall,
        "iterations": {"outer_mean": float(iters[:, 0].mean()), "inner_f_mean": float(iters[:, 1].mean()),
                       "inner_x_mean": float(iters[:, 2].mean()), "inner_total_max": int((iters[:, 1] + iters[:, 2]).max()),
                       "converged_frac": float((out["status"] == 0).mean())},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak if fp64_peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one solve_kernel launch of this workload, from
                     # the ncu capture named in profiles/r02_ncu_constants.json; algorithmic: 10.9 MB
                     "traffic": traffic,
                     "peak_source": "DFMA micro-benchmark measured in this run (bunmpc_measure_fp64_peak); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                     "kernel": "solve_kernel", "kernel_ms": 1e3 * k_time, "algorithmic_gflop_per_launch": flops * 1e-9},
        # the busiest memory pipe (ncu): shared-memory data stage; wavefronts per inner iteration from the ncu capture of
        # this same workload, scaled by the live iteration counters and the live kernel time
        "roofline_smem": smem_roofline(iters, k_time, clocks, info, ncu),
        **({"shard_balance": rank_info} if rank_info else {}),
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_ach / hbm_peak,
                         "traffic": traffic,   # same ncu capture as roofline.traffic
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
    }
    if others is not None:
        line["other_workloads"] = others
    # the CPU leg runs on rank 0 of a single-GPU run only (the other ranks of a multi-GPU run would just wait)
    if not args.no_cpu and world == 1:
        cores = cpu_cores()
        cb = synthetic.config(1, B=B, seed=0)
        v, n, dt = run_cpu(cb, cores, B, budget_s=20.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {n} instances of the same batch, {cores} threads, {dt:.1f} s, "
                                          "oracle/bicon_oracle.c (gcc -O3), solve only",
                                "single_solve_latency": cpu_single_solve_latency(cb)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
