/*
 * bicon_oracle.h -- CPU oracle for the BiConMP centroidal biconvex solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bunmpc_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (/root/reference/iterative_supervised_learning)
 * cannot be built here with its real linear algebra (Eigen 3 is absent from the
 * image and from the reference tree) and it ships no golden vectors or runnable
 * tests for this path.  This file restates the four
 * reference sources line by line; where Eigen's internal evaluation order is
 * not visible from the reference sources, the order is DEFINED here (see
 * "canonical evaluation order" in bicon_oracle.c) and the CUDA path reproduces
 * it bit for bit.  oracle/_ref (reference sources compiled against a minimal
 * stand-in header, see oracle/refshim/) cross-checks the control flow.
 *
 * Follows:
 *   src/dynamics/centroidal.cpp:6-127, include/dynamics/centroidal.hpp:22-27
 *   src/solvers/problem.cpp:11-56
 *   src/solvers/fista.cpp:6-70, include/solvers/fista.hpp:49-60
 *   src/motion_planner/biconvex.cpp:6-142, include/motion_planner/biconvex.hpp:148-160
 *   src/motion_planner/kino_dyn.cpp:83-99 (cold warm start, done by the caller)
 */
#ifndef BICON_ORACLE_H
#define BICON_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Solver constants; defaults are the reference's member initialisers
 * (biconvex.hpp:148-160, fista.hpp:52-60, biconvex.cpp:20-21). */
typedef struct {
    int    max_outer;   /* num_iters argument of BiConvexMP::optimize (100 cyclic / 50 acyclic) */
    int    max_inner;   /* maxit = 150 */
    double tol;         /* FISTA exit on ||prox-gradient|| : 1e-5 */
    double exit_tol;    /* outer exit on ||dyn violation||  : 1e-3 */
    double beta;        /* line-search growth : 1.5 */
    double mu;          /* friction coefficient : 1.0 (setter is not bound in python) */
    int    use_fma;     /* 0: separate multiply and add everywhere (x86-64 Release build of the
                           reference has no FMA); 1: the fused variant mirrored by the GPU's
                           BUNMPC_ARITH_FMA mode */
    int    reduction;   /* 0: the canonical order of dense sums (triples, then a tree over 32-blocks of the
                           triple sums -- rule (5) in bicon_oracle.c; what oracle/refshim and the GPU kernels
                           use); 32: a plain tree over 32-leaf blocks, a rounding variant for sensitivity tests */
    int    storage;     /* 0: ATA_ and A_ entries in binary64; 1: rounded to binary32 after set_data (arithmetic
                           stays binary64) -- the GPU's BUNMPC_ARITH_MIXED mode */
} bicon_params;

void bicon_default_params(bicon_params *p);

/* One instance.  All arrays are caller-owned, float64, C-contiguous.
 *   nx = 9*(n_col+1), nf = 3*n_eff*n_col                                  */
typedef struct {
    int n_col, n_eff;
    double m;                 /* robot mass */
    double rho;               /* penalty (set_rho) */
    const double *x_init;     /* [9] */
    const double *cnt_plan;   /* [n_col][n_eff][4] rows (c, x, y, z)  -- set_contact_plan */
    const double *dt;         /* [n_col] */
    const double *Qx;         /* [nx] diagonal of Q_x  (create_cost_X / set_cost_x) */
    const double *qx;         /* [nx] */
    const double *Qf;         /* [nf] diagonal of Q_f */
    const double *qf;         /* [nf] */
    const double *lbx;        /* [nx] */
    const double *ubx;        /* [nx] */
    const double *X0;         /* [nx] warm start (set_warm_start_vars) */
    const double *F0;         /* [nf] */
    const double *P0;         /* [nx] */
    double L_f, L_x;          /* FISTA step state carried by the object (fresh: 506.25, 2.25e6) */
} bicon_problem;

typedef struct {
    double *X;                /* [nx] */
    double *F;                /* [nf] */
    double *P;                /* [nx] */
    double L_f, L_x;          /* state after the solve */
    int    outer_iters;       /* outer iterations executed */
    int    inner_f, inner_x;  /* total FISTA iterations (calls of compute_step_length) */
    int    ls_f, ls_x;        /* line-search rejections (L *= beta) */
    double viol;              /* ||A_f X - b_f|| of the last outer iteration */
    int    status;            /* 0 converged, 1 max_outer reached, 2 NaN */
    double *viol_hist;        /* optional [max_outer], may be NULL (collect_statistics) */
} bicon_result;

/* Opaque per-(n_col, n_eff) workspace (symbolic patterns + work vectors). */
typedef struct bicon_ws bicon_ws;
bicon_ws *bicon_ws_create(int n_col, int n_eff);
void      bicon_ws_destroy(bicon_ws *ws);

/* BiConvexMP::optimize (biconvex.cpp:80-120). Returns 0, or -1 on bad arguments. */
int bicon_solve(bicon_ws *ws, const bicon_problem *p, const bicon_params *prm, bicon_result *out);

/* Host-side builders (biconvex.cpp:27-78). */
void bicon_create_bound_constraints(int n_col, int n_eff, const double *cnt_plan,
                                    const double *b /* [n_col][6] */,
                                    double *lbx, double *ubx);
void bicon_create_cost_X(int n_col, const double *W_X /* [9 n] */, const double *W_X_ter /* [9] */,
                         const double *X_ter /* [9] */, const double *X_nom /* [9 n] */,
                         double *Qx, double *qx);

/* Dense debug accessors (biconvex.hpp:30-51): A_x [nx][nf], b_x [nx], A_f [nx][nx], b_f [nx], row-major. */
void bicon_dense_x_mat(int n_col, int n_eff, double m, const double *cnt_plan, const double *dt,
                       const double *X, double *A_x, double *b_x);
void bicon_dense_f_mat(int n_col, int n_eff, double m, const double *cnt_plan, const double *dt,
                       const double *F, const double *x_init, double *A_f, double *b_f);

/* Batch driver used as the CPU baseline: B independent instances in struct-of-arrays form
 * (every array has a leading batch dimension, contiguous), n_threads worker threads over
 * disjoint contiguous instance ranges.  Outputs as in bicon_result, batched.
 * iters: [B][5] = outer, inner_f, inner_x, ls_f, ls_x.  Returns 0. */
int bicon_solve_batch(int B, int n_col, int n_eff, const double *m, const double *rho,
                      const double *x_init, const double *cnt_plan, const double *dt,
                      const double *Qx, const double *qx, const double *Qf, const double *qf,
                      const double *lbx, const double *ubx,
                      const double *X0, const double *F0, const double *P0,
                      const double *L_in /* [B][2] */, const bicon_params *prm, int n_threads,
                      double *X, double *F, double *P, double *L_out, int *iters,
                      double *viol, int *status);

#ifdef __cplusplus
}
#endif
#endif
