// Getters of the solve_kernel instantiations.  Each group lives in its own translation unit (solve_inst.cu compiled
// with -DINST_GROUP / -DINST_ARITH) so that the groups build in parallel; capi.cu only sees function pointers.
#pragma once
#include "kernels.cuh"

namespace bunmpc {

typedef void (*solve_fn)(const SolveArgs);

// group 0: horizons fixed at compile time (n = 20, 24, 30; nullptr for any other n)
// group 3: doubled horizons fixed at compile time (n = 40; 48 and 60 with combined roles)
// group 1: split warp roles, horizon read at run time
// group 2: combined warp roles (long horizons), horizon read at run time
// suffix: 0 = BUNMPC_ARITH_STRICT, 1 = BUNMPC_ARITH_FMA
solve_fn solve_inst_0_0(int n, int nthreads);
solve_fn solve_inst_0_1(int n, int nthreads);
solve_fn solve_inst_1_0(int n, int nthreads);
solve_fn solve_inst_1_1(int n, int nthreads);
solve_fn solve_inst_2_0(int n, int nthreads);
solve_fn solve_inst_2_1(int n, int nthreads);
solve_fn solve_inst_3_0(int n, int nthreads);
solve_fn solve_inst_3_1(int n, int nthreads);

}  // namespace bunmpc
