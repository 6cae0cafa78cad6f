// Getters of the solve_kernel instantiations.  Each (threads per CTA, arithmetic) pair lives in its own translation
// unit (solve_inst.cu compiled with -DINST_NT / -DINST_ARITH) so that they build in parallel; capi.cu only sees
// function pointers.
#pragma once
#include "kernels.cuh"

namespace bunmpc {

typedef void (*solve_fn)(const SolveArgs);

// Threads per CTA -> CTAs per SM the register budget is set for (__launch_bounds__): a thread keeps three Hessian
// rows in registers and wants ~168 of them, so an SM (64K registers) holds 384 solver threads: 4 CTAs of 96, 3 of 128, ...
// From 512 threads on the budget shrinks (128 / 80 / 64 registers) and the Hessian rows spill to local memory.
#define BUNMPC_NT_LIST(X) X(32, 12) X(64, 6) X(96, 4) X(128, 3) X(192, 2) X(256, 1) X(384, 1) X(512, 1) X(768, 1) X(1024, 1)
// experimental occupancy variants of the 96-thread kernel (BUNMPC_CTAS=5|6 in the environment, see capi.cu)
solve_fn solve_inst_x96(int arith, int ctas);

// smallest CTA size that holds a horizon: e*n force threads and 3(n+1) state/row threads
inline int solve_threads(int n, int e)
{
    const int need = (e * n > 3 * (n + 1)) ? e * n : 3 * (n + 1);
#define BUNMPC_PICK_NT(NT, MINB) if (need <= NT) return NT;
    BUNMPC_NT_LIST(BUNMPC_PICK_NT)
#undef BUNMPC_PICK_NT
    return 0;
}

// suffix: threads per CTA, then 0 = BUNMPC_ARITH_STRICT, 1 = BUNMPC_ARITH_FMA, 2 = BUNMPC_ARITH_MIXED
#define BUNMPC_DECL_INST(NT, MINB) solve_fn solve_inst_##NT##_0(); solve_fn solve_inst_##NT##_1(); solve_fn solve_inst_##NT##_2();
BUNMPC_NT_LIST(BUNMPC_DECL_INST)
#undef BUNMPC_DECL_INST

inline solve_fn solve_pick(int nthreads, int arith)
{
#define BUNMPC_PICK_FN(NT, MINB) if (nthreads == NT) return arith == 2 ? solve_inst_##NT##_2() : (arith ? solve_inst_##NT##_1() : solve_inst_##NT##_0());
    BUNMPC_NT_LIST(BUNMPC_PICK_FN)
#undef BUNMPC_PICK_FN
    return nullptr;
}

}  // namespace bunmpc
