#!/usr/bin/env python
"""Benchmark of the centroidal MPC hot path (BASELINE.json metric: centroidal MPC solves/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path (create_* + BiConvexMP::optimize for every instance) over one batch of
synthetic perturbed Solo12-trot states: BASELINE config[1], B = 1024 instances per GPU, horizon 20.
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


def smem_roofline(iters, k_time, clocks, info):
    wf_per_iter = 572.0
    mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz")
    if not mhz or not k_time:
        return None
    wf = wf_per_iter * float(iters[:, 1].sum() + iters[:, 2].sum())
    ach = wf / (k_time * mhz * 1e6 * info["num_sms"])
    return {"bound": "shared-memory data pipe", "achieved": ach, "peak": 1.0, "unit": "wavefronts/clk/SM", "frac": ach,
            "peak_source": "nominal 1 wavefront (128 B) per clock per SM",
            "wavefronts_per_inner_iteration": wf_per_iter,
            "ncu_pct_of_peak_sustained_elapsed": 57.8}   # l1tex__data_pipe_lsu_wavefronts_mem_shared, profiled launch


def other_workloads(device, arith):
    """Short single-GPU runs of the other BASELINE.json configs (per-GPU shard sizes), host-buffer path, one warm-up
    + one timed solve each: reported for context, not part of `value`."""
    from bunmpc_b200 import synthetic
    from bunmpc_b200.solver import BatchSolver
    out = {}
    cases = [("config2_go2_trot_B2048_n20 (16384/8 GPUs)", lambda: synthetic.config(2, B=2048)),
             ("config3_go2_bound_B512_n48 (reference algorithm diverges: cone quirk Q2, see DESIGN.md 6)", lambda: synthetic.config(3, B=512)),
             ("solo12_bound_B1024_n24 (bound gait, its own horizon)", lambda: synthetic.perturbed(1024, "solo12", "bound", seed=0)),
             ("solo12_jump_B1024_n30 (jump gait, its own horizon)", lambda: synthetic.perturbed(1024, "solo12", "jump", seed=0)),
             ("solo12_bound_B1024_n48 (config 3 gait and horizon, Solo12 mass)", lambda: synthetic.perturbed(1024, "solo12", "bound", seed=0, horizon_scale=2.0)),
             ("solo12_jump_B512_n60", lambda: synthetic.perturbed(512, "solo12", "jump", seed=0, horizon_scale=2.0)),
             ("config4_bayes_goal+weight_samples_B8192_n20 (65536/8 GPUs)", lambda: synthetic.config(4, B=8192)),
             ("solo12_trot_B8192_n20 (saturated)", lambda: synthetic.config(1, B=8192, seed=1))]
    for name, make in cases:
        b = make()
        s = BatchSolver(b.n_col, b.n_eff, max_batch=b.B, device=device)
        s.solve(b.select(np.arange(min(b.B, 256))), arith=arith)
        t0 = time.perf_counter()
        sol = s.solve(b, arith=arith)
        dt = time.perf_counter() - t0
        out[name] = {"solves_per_s_e2e": b.B / dt, "ms": 1e3 * dt, "n_col": b.n_col,
                     "outer_mean": float(sol.iters[:, 0].mean()), "inner_mean": float(sol.iters[:, 1:3].sum(1).mean()),
                     "converged_frac": float((sol.status == 0).mean()), "nan_frac": float((sol.status == 2).mean())}
        s.close()
    # f-1: only the centroidal states cross PCIe; contact plan and references are built on the device
    from bunmpc_b200.motions import GAITS, ROBOTS
    import torch
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
    out["solo12_trot_B1024_from_centroidal_states (device-side problem builder)"] = {
        "solves_per_s_e2e": B / dt, "ms": 1e3 * dt, "h2d_bytes": int(B * 8 * (9 + 12 + 1 + 3 + 1 + 2))}
    s.close()
    # f-3: lock-step rollouts, one batched solve per replanning tick, step sizes carried, host-side plan builder and plant
    from bunmpc_b200.rollout import EpisodeState, LockstepRollouts, TrackingPlant
    st = EpisodeState(com, vcom, amom, np.array(foot), t0s.astype(np.float64), np.zeros(B))
    roll = LockstepRollouts(rb, gp, plant=TrackingPlant(0.002, 0.02, 0.0, seed=1), device=device)
    roll.run(st, v_des, 0.0, n_ticks=1)
    t0 = time.perf_counter()
    rec = roll.run(st, v_des, 0.0, n_ticks=6)
    dt = time.perf_counter() - t0
    out["lockstep_rollouts_B1024_6_ticks (host plan builder + plant in the loop)"] = {
        "solves_per_s_e2e": int((np.array([(r > 0).any(1).sum() for r in rec.iters])).sum()) / dt, "ms": 1e3 * dt,
        "failed": int((rec.failed_at >= 0).sum())}
    return out


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
    ap.add_argument("--arith", default="strict", choices=["strict", "fma"])
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
    from bunmpc_b200 import synthetic, ARITH_FMA, ARITH_STRICT
    from bunmpc_b200.solver import BatchSolver

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    arith = ARITH_FMA if args.arith == "fma" else ARITH_STRICT
    B = args.batch

    # per-rank shard of the job: weak scaling, every GPU gets its own B perturbed states (seed = rank)
    batch = synthetic.config(1, B=B, seed=rank)
    solver = BatchSolver(batch.n_col, batch.n_eff, max_batch=B, device=local_rank)
    dev = solver.upload(batch)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    gathered = None
    if world > 1:   # the path's only exchange step: gather the solved trajectories (NCCL over NVLink)
        gathered = [torch.empty((world * B, solver.nf), dtype=torch.float64, device="cuda"),
                    torch.empty((world * B, solver.nx), dtype=torch.float64, device="cuda")]

    stats = torch.zeros(17, dtype=torch.float64, device="cuda")

    def step_resident():
        out = solver.solve_resident(dev, arith=arith)
        if world > 1:
            # the path's only exchange steps (BASELINE.json): gather the solved trajectories and all-reduce the
            # sufficient statistics of the Bayesian goal update (goal = desired velocity, error proxy = ||viol||)
            dist.all_gather_into_tensor(gathered[0], out["F"])
            dist.all_gather_into_tensor(gathered[1], out["X"])
            g = dev.fields["X_ter"][:, 3:6]
            e = torch.nan_to_num(out["viol"], nan=0.0)
            stats[0] = float(g.shape[0])
            stats[1:4] = g.sum(0)
            stats[4:13] = (g[:, :, None] * g[:, None, :]).sum(0).reshape(9)
            stats[13] = e.sum()
            stats[14:17] = (e[:, None] * g).sum(0)
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        return out

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
    iters = out["iters"]

    # ---- solve-kernel time alone (events around the solve launch only), for the roofline ----
    kt = []
    for _ in range(max(3, min(args.steps, 5))):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        solver.solve_resident(dev, arith=arith)
        b.record(stream)
        torch.cuda.synchronize()
        kt.append(a.elapsed_time(b) * 1e-3)
    k_time = float(np.mean(kt))

    # ---- e2e: host-buffer C ABI call, pinned inputs/outputs, copies inside the timed region ----
    import ctypes as C
    from bunmpc_b200 import _lib
    pinned = {}
    for f in _lib.COMPACT_FIELDS:
        a = getattr(batch, f)
        if a is None:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pinned[f] = t
        setattr(batch, f, t.numpy())
    outbuf = {k: torch.empty(v.shape, dtype=torch.from_numpy(v).dtype).pin_memory().numpy() for k, v in out.items()}
    for _ in range(2):
        solver.solve(batch, arith=arith, out=outbuf)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        sol = solver.solve(batch, arith=arith, out=outbuf)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    h2d = batch.input_bytes()
    d2h = int(sum(v.nbytes for v in outbuf.values()))

    # ---- p50 latency of a single solve through the host API (B = 1) ----
    one = batch.select(np.arange(1))
    lat = []
    for i in range(25):
        t0 = time.perf_counter()
        solver.solve(one, arith=arith)
        lat.append(time.perf_counter() - t0)
    p50_ms = float(np.median(lat[5:]) * 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the solve kernel ----
    fp64_peak = solver.measure_fp64_peak()
    flops = algorithmic_flops(batch.n_col, iters)
    achieved = flops / k_time * 1e-12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = algorithmic_bytes(batch.n_col, B) / k_time * 1e-9
    info = solver.kernel_info()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_gpu_per_step": B, "n_col": batch.n_col, "n_eff": batch.n_eff,
                   "arith": args.arith, "l2": "flushed between steps (256 MiB write)",
                   "sharding": "independent instances per rank, seed=rank" + (", nccl all_gather of F,X and all_reduce of 17 posterior statistics per step" if world > 1 else ""),
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
                     # dram__bytes_read.sum + dram__bytes_write.sum of one solve_kernel launch of this workload
                     # (ncu --set full, profiles/r01_solve_kernel_s2_ncu_summary.txt); algorithmic: 10.9 MB
                     "traffic": 13452544 if B == 1024 else None,
                     "peak_source": "DFMA micro-benchmark measured in this run (bunmpc_measure_fp64_peak); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                     "kernel": "solve_kernel", "kernel_ms": 1e3 * k_time, "algorithmic_gflop_per_launch": flops * 1e-9},
        # the busiest pipe (ncu): shared-memory data stage.  572 wavefronts per inner iteration (403 without bank
        # conflicts) is the ncu count of the profiled launch of this same workload (profiles/r01_solve_kernel_s2_
        # ncu_summary.txt: 57.8 % of peak); scaled here by the live iteration counters and the live kernel time.
        "roofline_smem": smem_roofline(iters, k_time, clocks, info),
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_ach / hbm_peak,
                         "traffic": 13452544 if B == 1024 else None,   # same ncu capture as roofline.traffic
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
    }
    # context legs run on rank 0 of a single-GPU run only (the other ranks of a multi-GPU run would just wait)
    if not args.no_extra and world == 1:
        line["other_workloads"] = other_workloads(local_rank, arith)
    if not args.no_cpu and world == 1:
        cores = cpu_cores()
        v, n, dt = run_cpu(synthetic.config(1, B=B, seed=0), cores, B, budget_s=25.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {n} instances of the same batch, {cores} threads, {dt:.1f} s, "
                                          "oracle/bicon_oracle.c (gcc -O3), solve only"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
