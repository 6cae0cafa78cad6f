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

#define BUNMPC_MINB_OF(NT, MINB) NT == INST_NT ? MINB:
constexpr int kMinBlocks = BUNMPC_NT_LIST(BUNMPC_MINB_OF) 1;

solve_fn INST_NAME() { return solve_kernel<4, INST_ARITH, INST_NT, kMinBlocks>; }

#if INST_NT == 96 && INST_ARITH == 2
solve_fn solve_inst_x96(int arith, int ctas)
{
    (void)arith;
    if (ctas == 5) return solve_kernel<4, 2, 96, 5>;
    if (ctas == 6) return solve_kernel<4, 2, 96, 6>;
    return nullptr;
}
#endif

}  // namespace bunmpc
