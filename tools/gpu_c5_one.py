"""One C5 batch solve on the default (tensor) large-state kernel, for ncu. usage: python tools/gpu_c5_one.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec = P.random_lqr_problem()
s = pkg.BatchSolver(spec, B)
X0 = P.normal_initial_states(spec, B)
for _ in range(2):
    s.set_inputs(X0)
    s.solve_al()
torch.cuda.synchronize()
print("ok", s.results()["status"][:4])
