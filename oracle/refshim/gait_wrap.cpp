// C wrapper around the UNMODIFIED reference class gait_planner::QuadrupedGait
// (/root/reference/iterative_supervised_learning/include/gait_planner/gait_planner.hpp), compiled against the Eigen
// stand-in of this directory into oracle/_ref/libgait_ref.so.  TEST INFRASTRUCTURE ONLY: it is what
// oracle/pinshim/gait_planner_cpp.py (the `GaitPlanner` the reference's gait generator imports) calls.
#include "gait_planner/gait_planner.hpp"

using gait_planner::QuadrupedGait;

extern "C" {

void *gait_create(double gait_period, const double *stance_percent, const double *phase_offset, int n_eff,
                  double step_height)
{
    Eigen::VectorXd sp(n_eff), po(n_eff);
    for (int i = 0; i < n_eff; ++i) { sp[i] = stance_percent[i]; po[i] = phase_offset[i]; }
    return new QuadrupedGait(gait_period, sp, po, step_height);
}
void gait_destroy(void *h) { delete static_cast<QuadrupedGait *>(h); }
int gait_get_phase(void *h, double t, int foot) { return static_cast<QuadrupedGait *>(h)->get_phase(t, foot); }
double gait_get_percent_in_phase(void *h, double t, int foot) { return static_cast<QuadrupedGait *>(h)->get_percent_in_phase(t, foot); }
double gait_get_phi(void *h, double t, int foot) { return static_cast<QuadrupedGait *>(h)->get_phi(t, foot); }

}  // extern "C"
