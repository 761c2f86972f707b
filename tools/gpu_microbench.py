"""Cycles per call of the per-knot device functions for one warp (run under gpurun)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
names = ["dfma", "sincos", "rk4_step", "knot_cost(k=1)", "knot_cost(k=0)", "quad_eval", "al_value(k=1)",
         "rk4_jacobian", "knot_expansion(k=1)", "riccati_step",
         "rollout/knot cold", "rollout/knot", "ro -cost", "ro -dyn", "ro -staging", "ro -stores", "ro -gdiv", "ro -all"]
for nm, spec in (("unicycle C2", P.unicycle_problem(P.K_THREE_OBSTACLES)),
                 ("triple C3", P.triple_integrator_problem(dof=2, N=50, add_constraints=True)),
                 ("cartpole C4", P.cartpole_problem(N=200))):
    s = pkg.BatchSolver(spec, 8); s.set_inputs(P.perturbed_initial_states(spec, 8, np.full(spec.n, 0.01)))
    out = (ctypes.c_longlong * 32)()
    rc = pkg.lib().altro_b200_microbench(s._h, out, 200)
    print(nm, "rc", rc, {n: int(out[i]) for i, n in enumerate(names)}, flush=True)
