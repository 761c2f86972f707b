"""Faithful vs skip_repeated_iterations at full C2 size: which instances differ, and how."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
B = 16384
X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
def solve(skip, engine="phased"):
    pkg.set_default_engine(engine)
    o = pkg.default_options(); o.skip_repeated_iterations = skip
    s = pkg.BatchSolver(spec, B, options=o); s.set_inputs(X0); s.solve_al()
    r = s.results(); X, U = s.trajectory()
    return r, X, U
a = solve(0); b = solve(1); c = solve(1, "fused"); d = solve(0)
for name, (r2, X2, U2) in (("skip phased", b), ("skip fused", c), ("faithful again", d)):
    r, X, U = a
    dX = np.abs(X - X2).reshape(B, -1).max(axis=1)
    bad = np.where(np.any(r["iters"] != r2["iters"], axis=1) | (r["cost"] != r2["cost"]) | (r["status"] != r2["status"]) | (dX > 0))[0]
    print(name, "differing:", len(bad), bad[:12].tolist())
    if len(bad):
        print("  iters faithful", r["iters"][bad[:6]].tolist(), "status", r["status"][bad[:6]].tolist())
        print("  iters other   ", r2["iters"][bad[:6]].tolist(), "status", r2["status"][bad[:6]].tolist())
        print("  dcost", (r["cost"][bad[:6]] - r2["cost"][bad[:6]]).tolist(), "dX", dX[bad[:6]].tolist())
