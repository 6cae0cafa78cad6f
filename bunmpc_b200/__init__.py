"""bunmpc_b200: B200-native batched BiConMP centroidal biconvex solve (one hot path of Atarilab/BUNMPC).

    from bunmpc_b200 import BiconvexMP            # drop-in for biconvex_mpc_cpp.BiconvexMP
    from bunmpc_b200 import solve_batch           # thousands of independent instances per launch

The kernels live in bunmpc_b200/csrc/libbunmpc.so (built in-tree, sm_100a); nothing here falls back to
the CPU.  Host-side modules (problem containers, gait planner, plan builder) import without the library.
"""
from .problem import BatchSolution, CentroidalBatch, SolverParams, L0_F, L0_X          # noqa: F401
from .gait_planner import GaitPlanner, QuadrupedGait                                    # noqa: F401
from .biconvex import BiconvexMP, BiConvexMP, CentroidalDynamics                        # noqa: F401
from .solver import BatchSolver, get_solver, solve_batch                                # noqa: F401
from .gait_gen import AbstractGaitGen, CyclicQuadrupedGaitGen, SoloMpcGaitGen            # noqa: F401
from .acyclic import ACYCLIC_MOTIONS, ACyclicMotionParams, SoloAcyclicGen                   # noqa: F401
from ._lib import ARITH_FMA, ARITH_MIXED, ARITH_STRICT, BunmpcError                                  # noqa: F401

__all__ = ["BiconvexMP", "BiConvexMP", "CentroidalDynamics", "BatchSolver", "solve_batch", "get_solver",
           "CentroidalBatch", "BatchSolution", "SolverParams", "GaitPlanner", "QuadrupedGait",
           "CyclicQuadrupedGaitGen", "SoloMpcGaitGen", "AbstractGaitGen", "SoloAcyclicGen", "ACyclicMotionParams",
           "ACYCLIC_MOTIONS",
           "ARITH_STRICT", "ARITH_FMA", "ARITH_MIXED", "BunmpcError"]
