// Getters of the solve_kernel instantiations.  Each (threads per CTA, arithmetic) pair lives in its own translation
// unit (solve_inst.cu compiled with -DINST_NT / -DINST_ARITH) so that they build in parallel; capi.cu only sees
// function pointers.
#pragma once
#include "kernels.cuh"

namespace bunmpc {

typedef void (*solve_fn)(const SolveArgs);

// Threads per CTA -> registers per thread (__maxnreg__).  A CTA is the worker warps that own the variables plus one
// service warp (kernels.cuh).  The pipelined FISTA loops keep the Hessian rows, the constraint rows, the iterate triples
// and the sums in flight in registers and want ~250 of them; a scheduler (16K registers) then holds two warps, an SM
// eight: 2 CTAs of 128 threads (the trot horizon), 1 of 256.  From 384 threads on the budget shrinks
// (168 / 128 / 96 / 80 / 64 registers) and rows move to shared-memory records or spill.
#define BUNMPC_NT_LIST(X) X(64, 255) X(96, 255) X(128, 255) X(160, 255) X(192, 255) X(256, 255) X(384, 168) X(512, 128) X(640, 96) X(768, 80) X(1024, 64)
// occupancy variant of the 128-thread kernel (BUNMPC_CTAS=3 in the environment, see capi.cu)
solve_fn solve_inst_x128_0(int ctas);
solve_fn solve_inst_x128_1(int ctas);
solve_fn solve_inst_x128_2(int ctas);
inline solve_fn solve_inst_x128(int arith, int ctas)
{
    return arith == 2 ? solve_inst_x128_2(ctas) : (arith ? solve_inst_x128_1(ctas) : solve_inst_x128_0(ctas));
}

// CTA size for a horizon: e*n force threads and 3(n+1) state/row threads in whole worker warps, plus one more warp --
// the service warp of short horizons (constraint rows, kernels.cuh), or at longer horizons a warp that has nothing to do
// but the line-search decisions -- unless that warp would cost residency or registers (128 -> 160 threads: one CTA per SM
// instead of two; 256 -> 288: 168 registers instead of 255; ...)
inline int solve_threads(int n, int e)
{
    const int work = (e * n > 3 * (n + 1)) ? e * n : 3 * (n + 1);
    const bool service = 3 * (n + 1) <= 64;
    const int w32 = 32 * ((work + 31) / 32);
    auto cls = [](int nt) { return nt <= 128 ? 0 : nt <= 256 ? 1 : nt <= 384 ? 2 : nt <= 512 ? 3 : nt <= 640 ? 4 : nt <= 768 ? 5 : 6; };
    const int need = (service || cls(w32 + 32) == cls(w32)) ? w32 + 32 : w32;
#define BUNMPC_PICK_NT(NT, MAXREG) if (need <= NT) return NT;
    BUNMPC_NT_LIST(BUNMPC_PICK_NT)
#undef BUNMPC_PICK_NT
    return 0;
}

// suffix: threads per CTA, then 0 = BUNMPC_ARITH_STRICT, 1 = BUNMPC_ARITH_FMA, 2 = BUNMPC_ARITH_MIXED
#define BUNMPC_DECL_INST(NT, MAXREG) solve_fn solve_inst_##NT##_0(); solve_fn solve_inst_##NT##_1(); solve_fn solve_inst_##NT##_2();
BUNMPC_NT_LIST(BUNMPC_DECL_INST)
#undef BUNMPC_DECL_INST

inline solve_fn solve_pick(int nthreads, int arith)
{
#define BUNMPC_PICK_FN(NT, MAXREG) if (nthreads == NT) return arith == 2 ? solve_inst_##NT##_2() : (arith ? solve_inst_##NT##_1() : solve_inst_##NT##_0());
    BUNMPC_NT_LIST(BUNMPC_PICK_FN)
#undef BUNMPC_PICK_FN
    return nullptr;
}

}  // namespace bunmpc
