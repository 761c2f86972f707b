"""Generates tests/golden/*.npz with the CPU oracle (which itself is pinned to the reference's
known-answer tests, tests/test_oracle_golden.py).  The fixtures are oracle outputs; for c1, c2 and c3 they are
also, number for number, the outputs of the reference's own solver sources compiled on the Eigen stand-in
(oracle/build_ref.py; checked by tests/test_oracle_vs_reference_build.py).  They let the GPU tests check the
CUDA path against committed vectors without running the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from altro_cpp_b200 import problems as P  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, spec, X0, use_al=True):
    out = ob.solve_batch(spec, X0, use_al=use_al, nthreads=os.cpu_count() or 1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X0=X0, X=out["X"], U=out["U"], K=out["K"], d=out["d"],
                        cost=out["cost"], viol=out["viol"], status=out["status"], iters=out["iters"])
    print(name, "instances", X0.shape[0], "status", np.bincount(out["status"]), "mean iterations", out["iters"][:, 2].mean())


if __name__ == "__main__":
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    dump("c2_unicycle_three_obstacles_al", spec, P.perturbed_initial_states(spec, 24, P.UNICYCLE_X0_SCALE))
    spec = P.unicycle_problem(P.K_TURN90)
    dump("c1_unicycle_turn90_ilqr", spec, P.perturbed_initial_states(spec, 16, P.UNICYCLE_X0_SCALE), use_al=False)
    spec = P.triple_integrator_problem(dof=2, N=50, add_constraints=True)
    dump("c3_triple_integrator_al", spec, P.perturbed_initial_states(spec, 16, P.TRIPLE_INTEGRATOR_X0_SCALE))
    spec = P.random_lqr_problem()
    dump("c5_random_lqr_al", spec, P.normal_initial_states(spec, 4))
