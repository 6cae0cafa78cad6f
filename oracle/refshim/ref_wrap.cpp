// C wrapper around the UNMODIFIED reference class motion_planner::BiConvexMP
// (/root/reference/iterative_supervised_learning/include/motion_planner/biconvex.hpp), compiled against the
// Eigen stand-in of this directory.  TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libbicon_ref.so, which is
// git-ignored.  The call protocol is the one the reference's gait generator uses
// (examples/mpc/abstract_cyclic_gen.py:391,611-614,663 and src/motion_planner/kino_dyn.cpp:83-99).
#define private public      // the test harness reads the FISTA step sizes and counts gradient products
#include "motion_planner/biconvex.hpp"
#undef private

#include <cmath>
#include <cstring>

using motion_planner::BiConvexMP;

static Eigen::VectorXd vec(const double *p, int n)
{
    Eigen::VectorXd v(n);
    for (int i = 0; i < n; ++i) v[i] = p[i];
    return v;
}

extern "C" {

void *ref_create(double m, int n_col, int n_eff) { return new BiConvexMP(m, n_col, n_eff); }
void ref_destroy(void *h) { delete static_cast<BiConvexMP *>(h); }

// one full solve through the reference's own setters; L_io = (L_f, L_x) in/out (fresh object: 506.25, 2.25e6)
// iters = outer, inner F, inner X, line-search rejections F, X
int ref_solve(void *h, int n_col, int n_eff, double rho, const double *x_init, const double *cnt_plan,
              const double *dt, const double *bounds, const double *W_X, const double *W_X_ter,
              const double *X_ter, const double *X_nom, const double *W_F, const double *X0, const double *F0,
              const double *P0, int num_iters, double *L_io, double *X, double *F, double *P, int *iters,
              double *viol)
{
    BiConvexMP &mp = *static_cast<BiConvexMP *>(h);
    const int n = n_col, e = n_eff, nx = 9 * (n + 1), nf = 3 * e * n;
    Eigen::shim::layout().n = n;                      // lets the stand-in pick the reduction layout by vector length
    Eigen::shim::layout().e = e;
    mp.set_rho(rho);
    for (int i = 0; i < n; ++i) {                      // abstract_cyclic_gen.py:391
        Eigen::MatrixXd cp(e, 4);
        for (int j = 0; j < e; ++j) for (int k = 0; k < 4; ++k) cp(j, k) = cnt_plan[4 * (e * i + j) + k];
        mp.set_contact_plan(cp, dt[i]);
    }
    Eigen::MatrixXd b(n, 6);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) b(i, k) = bounds[6 * i + k];
    mp.create_bound_constraints(b, 15.0, 15.0, 15.0);  // abstract_cyclic_gen.py:95-97,612
    mp.create_cost_X(vec(W_X, 9 * n), vec(W_X_ter, 9), vec(X_ter, 9), vec(X_nom, 9 * n));
    mp.create_cost_F(vec(W_F, nf));
    mp.set_warm_start_vars(vec(X0, nx), vec(F0, nf), vec(P0, nx));      // kino_dyn.cpp:98
    mp.fista_f.L_ = L_io[0];
    mp.fista_x.L_ = L_io[1];
    mp.log_statistics = true;
    mp.dyn_violation_hist_.clear();
    auto &cnt = Eigen::shim::product_counts();
    cnt.clear();
    mp.optimize(vec(x_init, 9), num_iters);            // kino_dyn.cpp:47
    const Eigen::VectorXd xo = mp.return_opt_x(), fo = mp.return_opt_f(), po = mp.return_opt_p();
    for (int i = 0; i < nx; ++i) { X[i] = xo[i]; P[i] = po[i]; }
    for (int i = 0; i < nf; ++i) F[i] = fo[i];
    const int ls_f = (int)std::lround(std::log(mp.fista_f.L_ / L_io[0]) / std::log(1.5));
    const int ls_x = (int)std::lround(std::log(mp.fista_x.L_ / L_io[1]) / std::log(1.5));
    L_io[0] = mp.fista_f.L_;
    L_io[1] = mp.fista_x.L_;
    iters[0] = (int)mp.dyn_violation_hist_.size();
    // compute_grad_obj multiplies ATA_ once per inner iteration (problem.cpp:54-56)
    iters[1] = (int)cnt[static_cast<const void *>(&mp.prob_data_f.ATA_)];
    iters[2] = (int)cnt[static_cast<const void *>(&mp.prob_data_x.ATA_)];
    iters[3] = ls_f;
    iters[4] = ls_x;
    *viol = iters[0] ? mp.dyn_violation_hist_.back() : 0.0;
    return 0;
}

}  // extern "C"
