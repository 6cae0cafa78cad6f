// One group of solve_kernel instantiations (see solve_inst.hpp).  Thread-count classes carry their own register
// budgets: the register file is split over 4 schedulers of 16K registers, so 16 warps allow 128 registers per
// thread, 20 warps 96, 24 warps 80, 32 warps 64.
#define BUNMPC_SOLVE_ONLY
#include "solve_inst.hpp"

#ifndef INST_GROUP
#error "compile with -DINST_GROUP=0|1|2|3 -DINST_ARITH=0|1"
#endif
#define INST_CAT2(a, b, c) a##b##_##c
#define INST_CAT(a, b, c) INST_CAT2(a, b, c)
#define INST_NAME INST_CAT(solve_inst_, INST_GROUP, INST_ARITH)

namespace bunmpc {

solve_fn INST_NAME(int n, int nthreads)
{
    constexpr int NE = 4, ARITH = INST_ARITH;
#if INST_GROUP == 0
    if (n == 20 && nthreads <= 480) return solve_kernel<NE, ARITH, 20, false, 480, 128>;   // BASELINE trot horizon
    if (n == 24 && nthreads <= 640) return solve_kernel<NE, ARITH, 24, false, 640, 96>;    // bound gait horizon (solo12_bound.py)
    if (n == 30 && nthreads <= 768) return solve_kernel<NE, ARITH, 30, false, 768, 80>;    // jump gait horizon (solo12_jump.py)
    return nullptr;
#elif INST_GROUP == 3
    // doubled horizons of the three gaits (analysis/solve_times_test.py:60-66); 48 and 60 run with combined roles
    if (n == 40 && nthreads <= 1024) return solve_kernel<NE, ARITH, 40, false, 1024, 64>;
    if (n == 48 && nthreads <= 768) return solve_kernel<NE, ARITH, 48, true, 768, 80>;
    if (n == 60 && nthreads <= 1024) return solve_kernel<NE, ARITH, 60, true, 1024, 64>;
    return nullptr;
#elif INST_GROUP == 1
    (void)n;
    if (nthreads <= 512) return solve_kernel<NE, ARITH, 0, false, 512, 128>;
    if (nthreads <= 640) return solve_kernel<NE, ARITH, 0, false, 640, 96>;
    if (nthreads <= 768) return solve_kernel<NE, ARITH, 0, false, 768, 80>;
    return solve_kernel<NE, ARITH, 0, false, 1024, 64>;
#else
    (void)n;
    if (nthreads <= 768) return solve_kernel<NE, ARITH, 0, true, 768, 80>;
    return solve_kernel<NE, ARITH, 0, true, 1024, 64>;
#endif
}

}  // namespace bunmpc
