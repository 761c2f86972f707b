"""One C2 solve (for ncu launch lists): python tools/gpu_one_solve.py [engine] [B] [skip]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
eng = sys.argv[1] if len(sys.argv) > 1 else "phased"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wl = sys.argv[4] if len(sys.argv) > 4 else "c2"
pkg.set_default_engine(eng)
if wl == "c2":
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES); scale = P.UNICYCLE_X0_SCALE
elif wl == "c3":
    spec = P.triple_integrator_problem(dof=2, N=50, add_constraints=True); scale = P.TRIPLE_INTEGRATOR_X0_SCALE
else:
    spec = P.cartpole_problem(N=200); scale = P.CARTPOLE_X0_SCALE
X0 = P.perturbed_initial_states(spec, B, scale)
o = pkg.default_options(); o.skip_repeated_iterations = skip
s = pkg.BatchSolver(spec, B, options=o)
s.set_inputs(X0); s.solve_al(); torch.cuda.synchronize()
print("done", s.engine, s.kernel_launches())
