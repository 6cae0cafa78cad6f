"""Lock-step batched rollouts and the goal posterior.

The reference runs its episodes strictly one after the other: `DataCollection.run`
(iterative_algorithm/data_collection.py:181-277) loops over perturbed restarts and, inside each, `rollout_mpc`
(simulation.py:340-580) replans every `plan_freq` seconds -- 60 sequential solves per episode.
`LocoSafeDagger.run_unperturbed` (locosafedagger_modified.py:449-614) does the same per sampled goal and then updates a
grid posterior over goals.  Here the loops are turned inside out: all episodes advance together and every replanning
tick is ONE batched solve on the GPU; the FISTA step sizes of each episode are carried from tick to tick exactly as
the reference's solver objects carry them, and an episode whose forces turn NaN is marked failed
and dropped from the following ticks (simulation.py:513-516).

What moves the robot between two replans is not part of the hot path (the reference uses PyBullet and an inverse-
dynamics controller, both out of scope and absent here), so it is a callback: `plant(state, batch, sol, dt) ->
state`.  `TrackingPlant` is the stand-in used by the tests and the bench: the robot follows its plan for `plan_freq`
seconds, optionally with Gaussian disturbances.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import dist
from .motions import BiconvexMotionParams, RobotConstants
from .plan_builder import build_batch
from .problem import BatchSolution, CentroidalBatch
from .solver import get_solver


@dataclass
class EpisodeState:
    """Centroidal state of B episodes."""
    com: np.ndarray        # [B,3]
    vcom: np.ndarray       # [B,3]
    amom: np.ndarray       # [B,3]
    foot_pos: np.ndarray   # [B,4,3]
    t: np.ndarray          # [B] gait time of each episode
    yaw: np.ndarray        # [B]

    def select(self, idx) -> "EpisodeState":
        return EpisodeState(self.com[idx], self.vcom[idx], self.amom[idx], self.foot_pos[idx], self.t[idx], self.yaw[idx])

    def assign(self, idx, other: "EpisodeState"):
        self.com[idx], self.vcom[idx], self.amom[idx] = other.com, other.vcom, other.amom
        self.foot_pos[idx], self.t[idx], self.yaw[idx] = other.foot_pos, other.t, other.yaw


class TrackingPlant:
    """Stand-in for the simulator: after `dt` the robot is where its plan says (linear interpolation between the
    knots, like the 1 kHz interpolation of abstract_cyclic_gen.py:677-692), feet that are in contact at that time sit
    on their planned contact locations, plus optional Gaussian disturbances (the perturbations of
    data_collection.py:232-252 applied at the centroidal level)."""

    def __init__(self, sigma_com=0.0, sigma_vcom=0.0, sigma_amom=0.0, seed=0):
        self.sigma = (sigma_com, sigma_vcom, sigma_amom)
        self.rng = np.random.default_rng(seed)

    def __call__(self, state: EpisodeState, batch: CentroidalBatch, sol: BatchSolution, dt: float) -> EpisodeState:
        B, n = sol.X.shape[0], batch.n_col
        X = sol.X.reshape(B, n + 1, 9)
        kdt = np.broadcast_to(batch.dt, (B, n))
        tk = np.concatenate([np.zeros((B, 1)), np.cumsum(kdt, axis=1)], axis=1)          # knot times
        k = np.clip((tk <= dt).sum(1) - 1, 0, n - 1)                                      # segment holding t = dt
        r = np.arange(B)
        a = ((dt - tk[r, k]) / kdt[r, k])[:, None]
        x = (1 - a) * X[r, k] + a * X[r, k + 1]
        cp = np.broadcast_to(batch.cnt_plan, (B, n, batch.n_eff, 4))[r, np.minimum(k + 1, n - 1)]   # plan at the new time
        foot = np.where(cp[:, :, 0:1] > 0, cp[:, :, 1:4], state.foot_pos)
        com, vcom, amom = x[:, 0:3].copy(), x[:, 3:6].copy(), x[:, 6:9].copy()
        for arr, s in zip((com, vcom, amom), self.sigma):
            if s > 0:
                arr += self.rng.normal(0.0, s, arr.shape)
        return EpisodeState(com, vcom, amom, foot, state.t + dt, state.yaw)


@dataclass
class RolloutRecord:
    """What the data-collection loop keeps per episode (data_collection.py:255-277: states, plans, failure flag)."""
    com: list = field(default_factory=list)      # per tick [B,3] (NaN rows for failed episodes)
    vcom: list = field(default_factory=list)
    F0: list = field(default_factory=list)       # first-knot forces of each plan [B, 3 n_eff]
    iters: list = field(default_factory=list)    # [B,5] solver counters per tick
    failed_at: np.ndarray = None                 # [B] tick at which the episode failed, -1 = completed

    def tracking_error(self, v_des: np.ndarray, w: np.ndarray = (1.0, 1.0, 0.0)) -> np.ndarray:
        """Weighted squared velocity-tracking error per episode (the role of locosafedagger_modified.py:560-579);
        failed episodes get +inf."""
        v = np.stack(self.vcom, axis=1)                                     # [B, ticks, 3]
        e = (((v - np.asarray(v_des)[:, None, :]) ** 2) * np.asarray(w)).sum(2).mean(1)
        return np.where(self.failed_at >= 0, np.inf, e)


class LockstepRollouts:
    """B episodes advanced together; one batched solve per replanning tick."""

    def __init__(self, robot: RobotConstants, params: BiconvexMotionParams, plan_freq: float = 0.05,
                 plant: Optional[Callable] = None, device: int = 0, horizon: Optional[int] = None,
                 solve_fn: Optional[Callable[[CentroidalBatch], BatchSolution]] = None, builder: str = "host"):
        """builder = "host": contact plans and references are built with numpy (plan_builder.build_batch) and the whole
        problem crosses PCIe every tick (4.3 KB per episode); builder = "device": only the centroidal states go up
        (230 B per episode), the problem is built by build_problem_kernel and solved where it lies
        (BatchSolver.build_device + solve_resident); what comes back is what the plant / controller consumes (the
        plan X, F, the step sizes and counters, and -- for TrackingPlant -- the contact plan it steps along)."""
        if builder not in ("host", "device"):
            raise ValueError("builder must be 'host' or 'device'")
        if builder == "device" and solve_fn is not None:
            raise ValueError("an injected solve_fn works on host batches: use builder='host'")
        if builder == "device" and horizon is not None and horizon != params.horizon():
            raise ValueError("the device builder plans over the gait's own horizon")
        self.robot, self.params, self.plan_freq, self.device = robot, params, plan_freq, device
        self.plant = plant if plant is not None else TrackingPlant()
        self.horizon = horizon if horizon is not None else params.horizon()
        self._solve_fn = solve_fn                 # tests inject the oracle here; default: the GPU solver
        self.builder = builder
        self.launches = 0
        self.h2d_bytes = 0                        # bytes that crossed PCIe towards the GPU (problem data or states)

    def _tick_device(self, st: "EpisodeState", v_des, w_des, amom_des, L0):
        """One replanning tick on the device path: states up, build + solve on the GPU, plan down."""
        solver = get_solver(self.horizon, 4, st.com.shape[0], self.device)
        dev = solver.build_device(self.robot, self.params, st.com, st.vcom, st.amom, st.foot_pos, st.t, v_des, w_des,
                                  yaw=st.yaw, amom_des=amom_des, L0=L0)
        o = solver.solve_resident(dev)
        self.launches += 1
        B, n, e = dev.B, self.horizon, 4
        self.h2d_bytes += B * 8 * (9 + 3 * e + 1 + 3 + 1 + 2 + (3 if amom_des is not None else 0) + (2 if L0 is not None else 0))
        host = {k: v.cpu().numpy() for k, v in o.items()}
        sol = BatchSolution(X=host["X"], F=host["F"], P=host["P"], L=host["L"], iters=host["iters"], viol=host["viol"],
                            status=host["status"], m=np.array([self.robot.mass]))
        f = dev.fields
        shared = lambda k: f[k].cpu().numpy()
        batch = CentroidalBatch(n, e, m=shared("m").reshape(-1), rho=shared("rho").reshape(-1), x_init=shared("x_init"),
                                cnt_plan=shared("cnt_plan").reshape(-1, n, e, 4), dt=shared("dt"), W_X=shared("W_X"),
                                W_X_ter=shared("W_X_ter"), X_nom=shared("X_nom"), X_ter=shared("X_ter"),
                                W_F=shared("W_F"), bounds=shared("bounds").reshape(-1, n, 6), L0=shared("L0"))
        return batch, sol

    def _solve(self, batch: CentroidalBatch) -> BatchSolution:
        self.launches += 1
        if self._solve_fn is not None:
            return self._solve_fn(batch)
        return get_solver(batch.n_col, batch.n_eff, batch.B, self.device).solve(batch)

    def run(self, state: EpisodeState, v_des, w_des, n_ticks: int, amom_des=None) -> RolloutRecord:
        B = state.com.shape[0]
        v_des = np.broadcast_to(np.asarray(v_des, dtype=np.float64), (B, 3)).copy()
        w_des = np.broadcast_to(np.asarray(w_des, dtype=np.float64), (B,)).copy()
        state = state.select(np.arange(B))                                   # private copy
        L = None                                                             # fresh FISTA objects: L0 = (506.25, 2.25e6)
        alive = np.ones(B, dtype=bool)
        rec = RolloutRecord(failed_at=np.full(B, -1, dtype=np.int64))
        for tick in range(n_ticks):
            idx = np.flatnonzero(alive)
            if idx.size == 0:
                break
            st = state.select(idx)
            ad = amom_des if amom_des is None or np.ndim(amom_des) < 2 else np.asarray(amom_des)[idx]
            if self.builder == "device":
                batch, sol = self._tick_device(st, v_des[idx], w_des[idx], ad, None if L is None else L[idx])
            else:
                batch = build_batch(self.robot, self.params, st.com, st.vcom, st.amom, st.foot_pos, st.t, v_des[idx],
                                    w_des[idx], yaw=st.yaw, horizon=self.horizon, amom_des=ad,
                                    L0=None if L is None else L[idx])
                self.h2d_bytes += batch.input_bytes()
                sol = self._solve(batch)
            if L is None:
                L = np.tile(sol.L[:1] * 0.0, (B, 1))
            L[idx] = sol.L                                                   # step sizes persist across replans (Q3)
            bad = np.isnan(sol.F).any(1) | (sol.status == 2)                 # simulation.py:513-516
            nxt = self.plant(st, batch, sol, self.plan_freq)
            state.assign(idx, nxt)
            com = np.full((B, 3), np.nan); vc = np.full((B, 3), np.nan)
            F0 = np.full((B, 3 * batch.n_eff), np.nan); it = np.zeros((B, 5), dtype=np.int64)
            com[idx], vc[idx], F0[idx], it[idx] = nxt.com, nxt.vcom, sol.F[:, : 3 * batch.n_eff], sol.iters
            com[idx[bad]] = np.nan; vc[idx[bad]] = np.nan
            rec.com.append(com); rec.vcom.append(vc); rec.F0.append(F0); rec.iters.append(it)
            rec.failed_at[idx[bad]] = tick
            alive[idx[bad]] = False
        return rec


class GoalPosterior:
    """Grid posterior over velocity goals (vx, vy, w): uniform prior on a 100^3 grid over [0,.3]x[-.1,.1]x[-.1,.1]
    (locosafedagger_modified.py:456-466), goals sampled from it (:404-423), Gaussian likelihood centred at an observed
    goal with sigma 0.1 (:357-384), posterior ~ prior * likelihood (:386-402).  `update_batch` folds in many observed
    goals at once (the product of their likelihoods), which is what one tick of batched rollouts produces; with
    several ranks the per-rank log-likelihood grids are summed with one all-reduce."""

    def __init__(self, n=100, vx=(0.0, 0.3), vy=(-0.1, 0.1), w=(-0.1, 0.1), sigma=0.1, device=None):
        self.axes = (np.linspace(*vx, n), np.linspace(*vy, n), np.linspace(*w, n))
        self.sigma = sigma
        self.device = device                                   # a torch device: the grid then lives in HBM
        self.p = np.full((n, n, n), 1.0 / n ** 3)
        if device is not None:
            import torch
            self._t = torch
            self.p = torch.full((n, n, n), 1.0 / n ** 3, dtype=torch.float64, device=device)
            self._ax = [torch.from_numpy(a).to(device) for a in self.axes]

    def sample(self, n_goals: int, rng: np.random.Generator) -> np.ndarray:
        """Draw goals (vx, vy, w) from the current posterior: [n_goals, 3]."""
        p = self.p.cpu().numpy() if self.device is not None else self.p
        flat = rng.choice(p.size, size=n_goals, p=p.ravel() / p.sum())
        i, j, k = np.unravel_index(flat, p.shape)
        return np.stack([self.axes[0][i], self.axes[1][j], self.axes[2][k]], axis=1)

    def sample_device(self, n_goals: int, generator=None):
        """Goals drawn from the posterior without leaving the GPU (torch.multinomial on the flattened grid): a
        [n_goals, 3] float64 tensor on the posterior's device.  (sample() copies the 100^3 grid to the host on every
        call; its numpy stream is what the CPU tests pin.)"""
        if self.device is None:
            raise ValueError("sample_device needs a posterior created with device=...")
        t = self._t
        flat = t.multinomial(self.p.reshape(-1), n_goals, replacement=True, generator=generator)
        n1, n2 = self.p.shape[1], self.p.shape[2]
        i, j, k = flat // (n1 * n2), (flat // n2) % n1, flat % n2
        return t.stack([self._ax[0][i], self._ax[1][j], self._ax[2][k]], dim=1)

    def log_likelihood(self, goals: np.ndarray, weights: Optional[np.ndarray] = None):
        """Sum over the observed goals of the log of the separable Gaussian likelihood on the grid."""
        g = np.asarray(goals, dtype=np.float64).reshape(-1, 3)
        wts = np.ones(len(g)) if weights is None else np.asarray(weights, dtype=np.float64)
        if self.device is not None:
            t = self._t
            gt, wt = t.from_numpy(g).to(self.device), t.from_numpy(wts).to(self.device)
            parts = [(-0.5 * ((self._ax[d][None, :] - gt[:, d:d + 1]) / self.sigma) ** 2 * wt[:, None]).sum(0) for d in range(3)]
            return parts[0][:, None, None] + parts[1][None, :, None] + parts[2][None, None, :]
        parts = [(-0.5 * ((self.axes[d][None, :] - g[:, d:d + 1]) / self.sigma) ** 2 * wts[:, None]).sum(0) for d in range(3)]
        return parts[0][:, None, None] + parts[1][None, :, None] + parts[2][None, None, :]

    def update_batch(self, goals: np.ndarray, weights: Optional[np.ndarray] = None, all_reduce: bool = False):
        ll = self.log_likelihood(goals, weights)
        if all_reduce:
            d = dist._dist()
            if d.is_initialized() and d.get_world_size() > 1:
                if self.device is None:
                    import torch
                    tl = torch.from_numpy(np.ascontiguousarray(ll))
                    d.all_reduce(tl)
                    ll = tl.numpy()
                else:
                    d.all_reduce(ll)
        if self.device is not None:
            post = self.p * self._t.exp(ll - ll.max())
            s = post.sum()
            self.p = post / s if float(s) > 0 else self._t.full_like(post, 1.0 / post.numel())
        else:
            self.p = dist.posterior_update(self.p, np.exp(ll - ll.max()))
        return self.p

    def update(self, observed_goal):
        """One observation, locosafedagger_modified.py:357-402 as its signature and docstring describe it: the Gaussian is
        centred at the observed goal (vx, vy, w)."""
        return self.update_batch(np.asarray(observed_goal, dtype=np.float64)[None])

    def update_as_called(self, v_des, w_des, error):
        """One observation the way the reference's ONLY call site passes it (locosafedagger_modified.py:611):
        `compute_likelihood(vx_vals, vy_vals, w_vals, v_des[0], v_des[1], w_des, self.errors[-1], sigma=0.1)` binds
        observed_goal = v_des[0] (unused), vx_obs = v_des[1], vy_obs = w_des, w_obs = error -- the arguments are shifted by
        one against the signature at :357, so the Gaussian the reference actually multiplies in is centred at
        (v_des[1], w_des, error).  A reference bug; `update` is the documented behaviour, this one the
        executed behaviour, and a caller reproducing the reference's runs needs this one."""
        v = np.asarray(v_des, dtype=np.float64).reshape(-1)
        return self.update_batch(np.array([[v[1], float(w_des), float(error)]], dtype=np.float64))
