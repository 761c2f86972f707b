import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
from oracle import binding as ob
spec = P.random_lqr_problem()
X0 = P.normal_initial_states(spec, 4)
for inner in (1, 2, 100):
    og = pkg.default_options(); og.max_iterations_inner = inner
    oo = ob.default_options(); oo.max_iterations_inner = inner
    s = pkg.BatchSolver(spec, 4, options=og); s.set_inputs(X0); s.solve_al()
    r = s.results(); Xg, Ug = s.trajectory(); Kg, dg = s.gains(); sc = s.scalars()
    ref = ob.solve_batch(spec, X0, options=oo, nthreads=4)
    print("inner", inner, "gpu iters", r["iters"].tolist(), "status", r["status"].tolist(), "ilqr", s.ilqr_status().tolist())
    print("          ref iters", ref["iters"].tolist(), "status", ref["status"].tolist())
    print("   cost", r["cost"], ref["cost"])
    print("   alpha", sc["alpha"], "z", sc["z"], "dJ", sc["dJ"], "grad", sc["grad"], "reg", sc["reg"], "dV", sc["dV0"], sc["dV1"], "init", sc["initial_cost"])
    print("   max|dX|", np.abs(Xg-ref["X"]).max(), "max|dU|", np.abs(Ug-ref["U"]).max(), "max|dK|", np.abs(Kg-ref["K"]).max(), "max|dd|", np.abs(dg-ref["d"]).max())
    o1 = ob.OracleSolver(spec, True); o1.set_options(oo); o1.set_initial_state(X0[0]); o1.solve_al()
    print("   oracle inst0: alpha", o1.stat("alpha"), "z", o1.stat("z"), "dJ", o1.stat("cost_decrease"), "grad", o1.stat("gradient"), o1.scalars())
