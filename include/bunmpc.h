/*
 * bunmpc.h -- C ABI of the B200-native batched BiConMP centroidal biconvex solver.
 *
 * This is the drop-in boundary for ONE path of Atarilab/BUNMPC: BiConvexMP::optimize and what it
 * calls.  It replaces the pybind11 module `biconvex_mpc_cpp`
 * (iterative_supervised_learning/srcpy/motion_planner/biconvex.cpp:15-44) for that path: every entry
 * point below names the reference interface it stands in for.  Plain pointers and sizes only; no
 * C++/torch types.  All floating-point data is IEEE binary64 ("f64"), C-contiguous.
 *
 * Shapes:  n = n_col, e = n_eff, nx = 9 (n + 1), nf = 3 e n, B = batch.
 * A `bunmpc_in` with batch_stride 0 is shared by all instances of the batch.
 *
 * There is no CPU fallback: every solve runs the sm_100a kernels in libbunmpc.so.
 */
#ifndef BUNMPC_H
#define BUNMPC_H

#ifdef __cplusplus
extern "C" {
#endif

#define BUNMPC_VERSION 100

enum {
    BUNMPC_OK = 0,
    BUNMPC_ERR_ARG = 1,          /* bad argument (null pointer, batch > max_batch, ...) */
    BUNMPC_ERR_UNSUPPORTED = 2,  /* n_col / n_eff / option outside what the kernels are built for */
    BUNMPC_ERR_CUDA = 3          /* CUDA runtime error, see bunmpc_last_error() */
};

/* per-instance exit status (biconvex.cpp:106-114) */
enum { BUNMPC_CONVERGED = 0, BUNMPC_MAX_ITERS = 1, BUNMPC_NAN = 2 };

/* arithmetic variants; BUNMPC_ARITH_STRICT reproduces the oracle's unfused operation order bit for bit,
 * BUNMPC_ARITH_FMA fuses the multiply-adds of the mat-vecs (oracle: use_fma = 1),
 * BUNMPC_ARITH_MIXED keeps the entries of ATA_ = 2(Q + rho A^T A) and of A_ in binary32 once set_data has formed them
 * (every operation, the line search and the exit tests stay binary64; oracle: storage = 1): the "FP32 mode" of the
 * path, a tolerance mode (1e-3 relative on forces and CoM/momentum), bit-exact against its own oracle twin. */
enum { BUNMPC_ARITH_STRICT = 0, BUNMPC_ARITH_FMA = 1, BUNMPC_ARITH_MIXED = 2 };

typedef struct bunmpc_solver bunmpc_solver;   /* opaque: device, stream, staging buffers, work queue */

/* Solver constants.  Defaults = the reference's member initialisers, none of which python can change
 * (biconvex.hpp:148-160, fista.hpp:52-60; num_iters is the argument of BiConvexMP::optimize). */
typedef struct {
    int    max_outer;   /* 100 (cyclic gaits) / 50 (acyclic) */
    int    max_inner;   /* 150 */
    double tol;         /* 1e-5 */
    double exit_tol;    /* 1e-3 */
    double beta;        /* 1.5 */
    double mu;          /* 1.0 */
    int    arith;       /* BUNMPC_ARITH_* */
    int    slice_outer; /* scheduling only, results never depend on it: outer iterations an instance runs before it is
                         * parked and re-queued (time slicing).  0 = automatic (8 when the batch exceeds the resident
                         * CTAs, else off), < 0 = off.  Parked instances are resumed longest predicted remainder first
                         * (eight queues; the estimate extrapolates the decay of the dynamics violation over the slice);
                         * environment BUNMPC_LONG_INNER = the scale of the predicted remaining inner iterations that
                         * separate the queues (x 0.2 .. x 3.6; default 1000) -- a tuning knob, it never changes a result */
} bunmpc_params;

typedef struct {
    const double *ptr;
    long long     batch_stride;   /* in elements; 0 = one copy shared by the whole batch */
} bunmpc_in;

/* What the gait generator hands to one BiconvexMP object per solve, batched
 * (abstract_cyclic_gen.py:391,611-614,663). */
typedef struct {
    int batch;
    bunmpc_in m;          /* [B]            ctor argument m            biconvex.cpp:6 */
    bunmpc_in rho;        /* [B]            set_rho                    biconvex.hpp:62-64 */
    bunmpc_in x_init;     /* [B][9]         optimize(x_init, .)        biconvex.cpp:80-82 */
    bunmpc_in cnt_plan;   /* [B][n][e][4]   set_contact_plan rows (c,x,y,z)  centroidal.cpp:39-49 */
    bunmpc_in dt;         /* [B][n]         set_contact_plan dt */
    bunmpc_in W_X;        /* [B][9n]        create_cost_X              biconvex.cpp:60-72 */
    bunmpc_in W_X_ter;    /* [B][9] */
    bunmpc_in X_nom;      /* [B][9n] */
    bunmpc_in X_ter;      /* [B][9] */
    bunmpc_in W_F;        /* [B][nf]        create_cost_F              biconvex.cpp:74-78 */
    bunmpc_in bounds;     /* [B][n][6]      create_bound_constraints   biconvex.cpp:27-58 */
    bunmpc_in L0;         /* [B][2]         FISTA step state (L_f, L_x); fresh object (506.25, 2.25e6)  biconvex.cpp:20-21 */
    bunmpc_in X0, F0, P0; /* [B][nx],[B][nf],[B][nx]  set_warm_start_vars (biconvex.hpp:66-70);
                             ptr == NULL selects the cold start of kino_dyn.cpp:83-99: X = tile(x_init), F = 0, P = 0 */
} bunmpc_compact_problem;

/* Same, with costs and bounds already expanded (set_cost_x / set_cost_f / set_bounds_x with diagonal Q,
 * biconvex.hpp:54-60,72-73). */
typedef struct {
    int batch;
    bunmpc_in m, rho, x_init, cnt_plan, dt;
    bunmpc_in Qx, qx;     /* [B][nx] diagonal of Q_x, q_x */
    bunmpc_in Qf, qf;     /* [B][nf] */
    bunmpc_in lbx, ubx;   /* [B][nx] */
    bunmpc_in L0;
    bunmpc_in X0, F0, P0; /* all three required here */
} bunmpc_expanded_problem;

/* Results, contiguous with leading batch dimension.  Any pointer may be NULL (not written). */
typedef struct {
    double *X;          /* [B][nx]  return_opt_x   biconvex.hpp:112-114 */
    double *F;          /* [B][nf]  return_opt_f   biconvex.hpp:116-118 */
    double *P;          /* [B][nx]  return_opt_p   biconvex.hpp:120-122 */
    double *L;          /* [B][2]   (L_f, L_x) after the solve: the state a FISTA object carries to its next call */
    int    *iters;      /* [B][5]   outer, sum inner F, sum inner X, line-search rejections F, X */
    double *viol;       /* [B]      ||A_f X - b_f|| of the last outer iteration */
    int    *status;     /* [B]      BUNMPC_CONVERGED / MAX_ITERS / NAN */
    double *viol_hist;  /* [B][max_outer] return_dyn_viol_hist (collect_statistics), NaN-padded */
    long long *cycles;  /* [B]      SM clock cycles the instance's CTA spent on it (in-kernel solve latency) */
} bunmpc_solution;

int         bunmpc_version(void);
const char *bunmpc_last_error(void);
void        bunmpc_default_params(bunmpc_params *p);

/* BiConvexMP::BiConvexMP(m, n_col, n_eff) (biconvex.cpp:6-25): allocates everything a batch of up to
 * max_batch instances needs on `device` (staging buffers, the time-slicing scratch, a stream).  No allocation afterwards. */
int  bunmpc_create(bunmpc_solver **out, int device, int n_col, int n_eff, int max_batch);
void bunmpc_destroy(bunmpc_solver *s);

/* kernel launches issued by this solver since creation (for bench.py's gpu_launches) */
long long bunmpc_launch_count(const bunmpc_solver *s);
/* CUDA occupancy of the solve kernel for this solver: CTAs per SM, threads per CTA, dynamic smem bytes */
int  bunmpc_kernel_info(const bunmpc_solver *s, int *ctas_per_sm, int *threads, int *smem_bytes, int *num_sms);

/* ---- device-pointer entry points: asynchronous on `stream` (a cudaStream_t; NULL = CUDA's default stream).
 *      All pointers in the structs are device pointers on the solver's device.
 *      ONE solve in flight per handle: the work queue, the parked states of the time slicing and the expanded problem
 *      are scratch of the handle, so a second solve on the same handle has to be ordered after the first (same stream,
 *      or an event); use one handle per concurrent stream. ---- */

/* create_bound_constraints + create_cost_X + create_cost_F for a batch (biconvex.cpp:27-78):
 * writes Qx,qx,lbx,ubx [B][nx] and Qf,qf [B][nf]. */
int bunmpc_expand_device(bunmpc_solver *s, const bunmpc_compact_problem *p,
                         double *Qx, double *qx, double *Qf, double *qf, double *lbx, double *ubx, void *stream);
/* BiConvexMP::optimize on expanded inputs (biconvex.cpp:80-120). */
int bunmpc_solve_expanded_device(bunmpc_solver *s, const bunmpc_expanded_problem *p, const bunmpc_params *prm,
                                 const bunmpc_solution *out, void *stream);
/* create_* followed by optimize: the whole per-solve protocol of abstract_cyclic_gen.py:391,611-614,663. */
int bunmpc_solve_compact_device(bunmpc_solver *s, const bunmpc_compact_problem *p, const bunmpc_params *prm,
                                const bunmpc_solution *out, void *stream);

/* ---- batched problem builder on the device -------------------------------------------------
 * What SoloMpcGaitGen does in python before it calls the solver -- create_cnt_plan
 * (examples/mpc/abstract_cyclic_gen.py:159-414: Raibert contact plan, phase lookup of gait_planner.cpp:41-58,
 * dt[0] rule) and the dynamics part of create_costs (:564-614: X_nom, X_ter) -- for a whole batch, starting
 * from centroidal states (pinocchio stays on the caller's side). */
typedef struct {
    double gait_period, gait_dt, gait_horizon;   /* motions/cyclic/*.py */
    double stance_percent[4], phase_offset[4];
    double hip_offsets[4][2];                     /* abstract_cyclic_gen.py:51-69 (x, y) */
    double foot_size;                             /* :31 */
    double nom_ht;
    double ori_correction[3];
    double I_zz;                                  /* composite yaw inertia, used when w_des != 0 (:604) */
    double W_X[9], W_X_ter[9], W_F[12], rho;      /* used only when per-instance scalings are given */
    int    swing_rule;                            /* planned location of a foot in the second half of its swing:
                                                     0 = hip + yaw step + Raibert step (SoloMpcGaitGen,
                                                     abstract_cyclic_gen.py:351-355), 1 = hip + yaw step only
                                                     (AbstractGaitGen, abstract_cyclic_gen1.py:211-215) */
    int    reserved_;
} bunmpc_gait;

typedef struct {
    int batch;
    bunmpc_in com, vcom, amom;     /* [B][3] each: X_init = [com, hg_lin/m, hg_ang] (:567-571) */
    bunmpc_in foot_pos;            /* [B][4][3] current end-effector positions (rounded to 3 dp like :213) */
    bunmpc_in t;                   /* [B] time inside the gait (replanning phase) */
    bunmpc_in v_des;               /* [B][3] desired velocity, already in the local frame (:642-643) */
    bunmpc_in w_des;               /* [B] */
    bunmpc_in cs_yaw;              /* [B][2] cos and sin of the base yaw (R of :172-177); unused when hip_xy is given */
    bunmpc_in hip_xy;              /* [B][4][2] (R hip_offset_j)[0:2], the rotated hip offsets of :279,347.  The reference
                                      forms them with numpy.matmul, whose rounding depends on the BLAS numpy is linked
                                      with, so a caller that wants the reference's bits passes the products themselves;
                                      ptr NULL = cos(yaw) ox - sin(yaw) oy, sin(yaw) ox + cos(yaw) oy, unfused */
    bunmpc_in amom_des;            /* [B][3] log3(R_des R_q^T) (:616-627); ptr NULL = 0 */
    bunmpc_in scales;              /* [B][3] multipliers of (W_X and W_X_ter, W_F, rho); ptr NULL = none */
} bunmpc_states;

/* device pointers; writes x_init [B][9], cnt_plan [B][n][4][4], dt [B][n], X_nom [B][9n], X_ter [B][9] and, when
 * st->scales.ptr != NULL, W_X [B][9n], W_X_ter [B][9], W_F [B][12n], rho [B] (otherwise those four may be NULL). */
int bunmpc_build_problem_device(bunmpc_solver *s, const bunmpc_gait *g, const bunmpc_states *st, double *x_init,
                                double *cnt_plan, double *dt, double *X_nom, double *X_ter, double *W_X,
                                double *W_X_ter, double *W_F, double *rho, void *stream);

/* ---- the same for the ACYCLIC generator (SoloAcyclicGen, examples/mpc/abstract_acyclic_gen.py:74-190): a motion is
 * three time tables (examples/motions/weight_abstract.py:46-83), every knot of a replan at time t is looked up in them.
 * All pointers are DEVICE pointers; the tables are uploaded once per motion. */
typedef struct {
    int n_cnt, n_nom, n_box;      /* segments per table (each >= 1) */
    const double *dt_arr;         /* [n_col] */
    const double *cnt_plan;       /* [n_cnt][4][6]  c, x, y, z, t_start, t_end per foot */
    const double *X_nom;          /* [n_nom][11]    9 values, t_start, t_end */
    const double *bounds;         /* [n_box][8]     6 values, t_start, t_end */
    const double *X_ter;          /* [9] */
    double t0;                    /* time the motion started (update_motion_params, :42-54) */
} bunmpc_acyclic_motion;

/* x_init [B][9] (passed through to X_nom's first knot, :183), t [B] replanning instants; writes cnt_plan [B][n][4][4],
 * dt [B][n], X_nom [B][9n], X_ter [B][9], bounds [B][n][6].  Asynchronous on `stream`. */
int bunmpc_build_acyclic_device(bunmpc_solver *s, const bunmpc_acyclic_motion *m, int batch, const bunmpc_in *x_init,
                                const bunmpc_in *t, double *cnt_plan, double *dt, double *X_nom, double *X_ter,
                                double *bounds, void *stream);

/* ---- multi-GPU jobs with ONE fresh-instance counter ----------------------------------------------------------
 * The instances of the path are independent, so what a multi-GPU job needs is balance, not a collective: every rank holds
 * all B instances of the job in its HBM and the CTAs of all GPUs pull instance ids from one counter in the memory of one
 * GPU (system-scope atomics over NVLink / NVSwitch peer memory) -- the GPUs then finish together whatever the
 * instances cost, where a fixed split i -> rank i mod G waits for the slowest shard.  Parked instances (time slicing)
 * stay on the GPU that started them.  Every rank solves a SUBSET of the rows of its output buffers and leaves the
 * other rows untouched (zero them first; the ranks' buffers then combine by a bitwise-exact integer sum).
 * Protocol: the owner calls bunmpc_job_counter_create and sends the 64-byte CUDA IPC handle to the other ranks'
 * processes, which call bunmpc_job_counter_open; every rank calls bunmpc_set_job_counter on its solver; then all ranks
 * call the same solve entry point the same number of times with batch = B, with a collective of the job (the
 * exchange of the results) between two solves.  counters == NULL switches back to the local counter. */
int bunmpc_job_counter_create(int device, void **counters, unsigned char ipc_handle[64]);
int bunmpc_job_counter_open(int device, const unsigned char ipc_handle[64], void **counters);
int bunmpc_job_counter_release(void *counters, int owner);
int bunmpc_set_job_counter(bunmpc_solver *s, void *counters, int owner);

/* Fused exchange: result buffers that the other GPUs of the job store into.  Every rank allocates its result rows with
 * bunmpc_peer_buffer_create (device memory + 64-byte CUDA IPC handle), sends the handle to the other ranks' processes,
 * opens theirs with bunmpc_peer_buffer_open and registers, per peer, where that peer's X, F, L, viol, iters and status
 * rows live (bunmpc_set_peer_results; the other members of bunmpc_solution are ignored).  From then on the CTA that
 * finishes instance b stores row b into its own `out` buffers AND into every registered peer's (plain stores to peer
 * memory over NVLink / NVSwitch from the solve kernel's epilogue): when every rank's solve has finished -- any
 * collective or barrier of the job after the solve tells -- every rank holds every row, and no collective moves results.
 * A rank must not start the next solve of the job while a peer still reads the rows of this one (one more barrier).
 * n_peers == 0 switches the stores off.  At most 15 peers. */
int bunmpc_peer_buffer_create(int device, unsigned long long bytes, void **ptr, unsigned char ipc_handle[64]);
int bunmpc_peer_buffer_open(int device, const unsigned char ipc_handle[64], void **ptr);
int bunmpc_peer_buffer_release(void *ptr, int owner);
int bunmpc_set_peer_results(bunmpc_solver *s, int n_peers, const bunmpc_solution *peers);

/* Sufficient statistics of the Bayesian goal update over one rank's shard (the reference's grid posterior with a Gaussian
 * likelihood centred at the sampled goal, locosafedagger_modified.py:357-402): goals [B][3] (batch stride in elements,
 * e.g. the desired velocity X_ter + 3 with stride 9), errors [B] (NaN counts as 0) -> out17 = [N, sum g (3),
 * sum g g^T (9), sum e, sum e g (3)], device pointers, asynchronous on `stream`, fixed summation order.  The 17 doubles
 * are what the ranks of a multi-GPU job all-reduce (NCCL). */
int bunmpc_goal_stats_device(bunmpc_solver *s, int batch, const bunmpc_in *goals, const bunmpc_in *errors, double *out17,
                             void *stream);

/* ---- host-pointer entry points: copy in, solve, copy out, synchronise.  Pointers are host pointers
 *      (pinned memory makes the copies asynchronous).  This is what the python BiconvexMP calls. ---- */
int bunmpc_solve_compact_host(bunmpc_solver *s, const bunmpc_compact_problem *p, const bunmpc_params *prm,
                              const bunmpc_solution *out);
int bunmpc_solve_expanded_host(bunmpc_solver *s, const bunmpc_expanded_problem *p, const bunmpc_params *prm,
                               const bunmpc_solution *out);

/* return_A_x / return_b_x / return_A_f / return_b_f (biconvex.hpp:30-51) for ONE instance, dense row-major,
 * host pointers: A_x [nx][nf], b_x [nx], A_f [nx][nx], b_f [nx].  Any output may be NULL. */
int bunmpc_centroidal_mats_host(bunmpc_solver *s, double m, const double *cnt_plan, const double *dt,
                                const double *X, const double *F, const double *x_init,
                                double *A_x, double *b_x, double *A_f, double *b_f);

/* Measurement aid for bench.py: FP64 FMA throughput of this GPU (TFLOP/s, 2 flops per DFMA) from a
 * register-only DFMA micro-benchmark; the denominator of the FP64-pipe roofline. */
int bunmpc_measure_fp64_peak(bunmpc_solver *s, double *tflops);

/* Self-test of the hoisted-reciprocal division used for g/L in the FISTA step (kernels.cuh div_fast):
 * compares it with IEEE division on n_pairs pseudo-random operand pairs, returns the number of mismatches. */
int bunmpc_selftest_division(bunmpc_solver *s, long long n_pairs, unsigned long long seed, long long *mismatches);

/* pinned host memory helpers for callers without a CUDA runtime of their own */
void *bunmpc_host_alloc(unsigned long long bytes);
void  bunmpc_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
