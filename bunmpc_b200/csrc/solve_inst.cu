// One solve_kernel instantiation (see solve_inst.hpp): -DINST_NT=<threads per CTA> -DINST_ARITH=0|1.
#define BUNMPC_SOLVE_ONLY
#include "solve_inst.hpp"

#if !defined(INST_NT) || !defined(INST_ARITH)
#error "compile with -DINST_NT=<threads> -DINST_ARITH=0|1"
#endif
#define INST_CAT2(a, b, c) a##b##_##c
#define INST_CAT(a, b, c) INST_CAT2(a, b, c)
#define INST_NAME INST_CAT(solve_inst_, INST_NT, INST_ARITH)

namespace bunmpc {

#define BUNMPC_MAXREG_OF(NT, MAXREG) NT == INST_NT ? MAXREG:
constexpr int kMaxReg = BUNMPC_NT_LIST(BUNMPC_MAXREG_OF) 1;

solve_fn INST_NAME() { return solve_kernel<4, INST_ARITH, INST_NT, kMaxReg>; }

#if INST_NT == 128
// occupancy variant of the 128-thread kernel (BUNMPC_CTAS=3 in the environment, see capi.cu): 3 CTAs per SM at 168
// registers instead of 2 at 255
#define INST_X128_NAME INST_CAT(solve_inst_x, INST_NT, INST_ARITH)
solve_fn INST_X128_NAME(int ctas)
{
    if (ctas == 3) return solve_kernel<4, INST_ARITH, 128, 168>;
    return nullptr;
}
#endif

}  // namespace bunmpc
