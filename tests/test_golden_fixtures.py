"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py with the
oracle; for c1 / c2 / c3 they are also, number for number, what the reference's own solver sources produce —
tests/test_oracle_vs_reference_build.py asks oracle/_ref for the same instances).  CPU: the oracle still
reproduces them (guards against silent drift of the checker).  GPU: the CUDA path matches them — same
iteration path, trajectories to 1e-8."""
import glob
import os

import numpy as np
import pytest

from altro_cpp_b200 import problems as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "c2_unicycle_three_obstacles_al": (lambda: P.unicycle_problem(P.K_THREE_OBSTACLES), True),
    "c1_unicycle_turn90_ilqr": (lambda: P.unicycle_problem(P.K_TURN90), False),
    "c3_triple_integrator_al": (lambda: P.triple_integrator_problem(dof=2, N=50, add_constraints=True), True),
    "c5_random_lqr_al": (lambda: P.random_lqr_problem(), True),
}


def test_every_fixture_has_a_case():
    names = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "*.npz")))
    assert names == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_fixture(oracle, name):
    make, al = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out = oracle.solve_batch(make(), g["X0"], use_al=al, nthreads=4)
    for key in ("status", "iters"):
        assert np.array_equal(out[key], g[key]), key
    # bit-equal on the host that generated them; libm picks CPU-specific sin/cos kernels, so allow
    # last-bit drift on other hosts
    for key in ("X", "U", "cost"):
        assert np.abs(out[key] - g[key]).max() <= 1e-10 * max(1.0, np.abs(g[key]).max()), key


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_matches_fixture(name):
    import altro_cpp_b200 as pkg
    make, al = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    spec = make()
    s = pkg.BatchSolver(spec, g["X0"].shape[0], use_constraints=al)
    s.set_inputs(g["X0"])
    (s.solve_al if al else s.solve_ilqr)()
    r = s.results()
    X, U = s.trajectory()
    assert np.array_equal(r["iters"], g["iters"]) and np.array_equal(r["status"], g["status"])
    scale = lambda a: max(1.0, np.abs(a).max())
    assert np.abs(X - g["X"]).max() / scale(g["X"]) <= 1e-8
    assert np.abs(U - g["U"]).max() / scale(g["U"]) <= 1e-8
    assert np.abs(r["cost"] - g["cost"]).max() / scale(g["cost"]) <= 1e-9
