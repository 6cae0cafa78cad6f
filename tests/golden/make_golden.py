"""Generates tests/golden/centroidal_cases.npz.

Run in the build container (where /root/reference is mounted and oracle/_ref has been built by
`make -C oracle/refshim`):   python tests/golden/make_golden.py

For every case the file stores the compact inputs of the solve and five result sets:
  ref    : the reference's OWN sources (biconvex.cpp, centroidal.cpp, problem.cpp, fista.cpp compiled in place
           against oracle/refshim's Eigen stand-in), driven through its own setters and optimize()
  oc     : the oracle in the canonical evaluation order (must equal `ref` bit for bit) = what the GPU STRICT mode
           reproduces
  ocf    : same with fused multiply-adds = what the GPU FMA mode reproduces
  ocm    : same with ATA_ / A_ entries stored in binary32 = what the GPU MIXED mode reproduces
  o32    : the oracle with plain 32-leaf reduction blocks -- a rounding variant (tests/test_oracle_golden.py)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bunmpc_b200 import synthetic                      # noqa: E402
from bunmpc_b200.problem import CentroidalBatch       # noqa: E402
from oracle import oracle                              # noqa: E402

FIELDS = ("m", "rho", "x_init", "cnt_plan", "dt", "W_X", "W_X_ter", "X_nom", "X_ter", "W_F", "bounds", "L0",
          "X0", "F0", "P0")
RES = ("X", "F", "P", "L", "iters", "viol", "status")


def cases():
    yield "trot_nominal", synthetic.nominal("solo12", "trot")
    yield "trot_perturbed", synthetic.perturbed(3, "solo12", "trot", seed=0)
    yield "bound_perturbed", synthetic.perturbed(2, "solo12", "bound", seed=1)
    yield "jump_perturbed", synthetic.perturbed(2, "solo12", "jump", seed=2)
    yield "go2_backtracks", synthetic.perturbed(2, "go2", "trot", seed=3)          # cap, L_x backtracks, one NaN
    b = synthetic.perturbed(2, "solo12", "trot", seed=7)
    b.L0 = np.array([[1.0, 40.0]])
    yield "tiny_step_rejections", b
    # warm start: second solve of the same object (state X,F,P and step sizes carried over, kino_dyn cold start off)
    b = synthetic.perturbed(2, "solo12", "trot", seed=9)
    first = oracle.solve(b)
    yield "warm_start", b.with_state(X0=first["X"], F0=first["F"], P0=first["P"], L0=first["L"])


def main():
    out = {}
    names = []
    for name, b in cases():
        names.append(name)
        out[f"{name}/n_col"], out[f"{name}/n_eff"] = b.n_col, b.n_eff
        for f in FIELDS:
            v = getattr(b, f)
            if v is not None:
                out[f"{name}/in/{f}"] = v
        sets = {"o32": oracle.solve(b, params=oracle.default_params(reduction=32)),
                "oc": oracle.solve(b),
                "ocf": oracle.solve(b, params=oracle.default_params(use_fma=1)),
                "ocm": oracle.solve(b, params=oracle.default_params(storage=1))}
        if oracle.ref_available():
            sets["ref"] = oracle.ref_solve(b)
            same = all(np.array_equal(sets["ref"][k], sets["oc"][k], equal_nan=True) for k in RES if k != "status")
            print(f"{name:24s} ref == oc: {same}   iters {sets['ref']['iters'].tolist()}")
            assert same, name
        for sname, res in sets.items():
            for k in RES:
                out[f"{name}/{sname}/{k}"] = res[k]
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "centroidal_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def load(path=None):
    """-> list of (name, CentroidalBatch, {set: {key: array}})"""
    path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "centroidal_cases.npz")
    z = np.load(path)
    res = []
    for name in z["names"]:
        kw = {f: z[f"{name}/in/{f}"] for f in FIELDS if f"{name}/in/{f}" in z.files}
        b = CentroidalBatch(int(z[f"{name}/n_col"]), int(z[f"{name}/n_eff"]), **kw)
        sets = {}
        for sname in ("ref", "oc", "ocf", "ocm", "o32"):
            if f"{name}/{sname}/X" in z.files:
                sets[sname] = {k: z[f"{name}/{sname}/{k}"] for k in RES}
        res.append((str(name), b, sets))
    return res


if __name__ == "__main__":
    main()
