"""The oracle restatement (oracle/altro_oracle.cpp) against the REFERENCE's own solver code, compiled here from the
sources under /root/reference by oracle/build_ref.py (on this repo's Eigen stand-in — Eigen itself is absent — and
driven through oracle/ref_shim/ref_driver.cpp).  Same problem definitions (the reference's examples/problems/*.hpp on
one side, altro_cpp_b200/problems.py on the other), same perturbed initial states; compared per instance:

* verdict and iteration counts (inner of the last AL iteration, outer, total): exactly — this is the reference's
  control flow (line-search accept / reject, regularisation, dual and penalty updates, termination tests);
* cost, stored violation, states and controls: exactly as well.  The oracle was written to perform the reference's
  operations in the reference's order, and the stand-in evaluates Eigen's expressions in the plain textbook order, so
  the two agree bit for bit; against a build on the real Eigen (blocked / vectorised products) the last bits of long
  sums would differ, which is the 1e-12 the reference's one C++-generated golden cost shows (auglag_test.cpp:348).

CPU only.  Skipped where /root/reference is not mounted (the GPU box): oracle/_ref cannot be built there and nothing
on the GPU path needs it.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from altro_cpp_b200 import problems as P  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import build_ref  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    if not build_ref.available():
        pytest.skip("the reference sources are not mounted here")
    return ctypes.CDLL(build_ref.build())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ref_solve(lib, entry, first_arg, constrained, x0, n, m, N, options=None):
    """options: [constraint_tolerance, SetPenalty, initial_penalty, max total, max inner, max outer], < 0 = default"""
    X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    opt = None if options is None else np.ascontiguousarray(options, dtype=np.float64)
    got = getattr(lib, entry)(ctypes.c_int(first_arg), ctypes.c_int(int(constrained)), _ptr(x0), _ptr(opt), _ptr(X),
                              _ptr(U), _ptr(sc), _ptr(it))
    assert got == N
    return dict(X=X, U=U, cost=sc[0], viol=sc[1], max_penalty=sc[2], initial_cost=sc[3], status=int(it[0]),
                inner=int(it[1]), outer=int(it[2]), total=int(it[3]))


def compare(lib, spec, entry, first_arg, constrained, X0, options=None, oracle_options=None):
    o = ob.solve_batch(spec, X0, options=oracle_options, use_al=constrained, nthreads=8, want_gains=False)
    same_path, worst = 0, dict(cost=0.0, X=0.0, U=0.0, viol=0.0)
    for b in range(X0.shape[0]):
        r = ref_solve(lib, entry, first_arg, constrained, X0[b], spec.n, spec.m, spec.N, options)
        mine = (int(o["status"][b]), int(o["iters"][b, 0]), int(o["iters"][b, 1]), int(o["iters"][b, 2]))
        theirs = (r["status"], r["inner"], r["outer"], r["total"])
        if mine != theirs:
            continue
        same_path += 1
        worst["cost"] = max(worst["cost"], abs(o["cost"][b] - r["cost"]) / max(1.0, abs(r["cost"])))
        worst["viol"] = max(worst["viol"], abs(o["viol"][b] - r["viol"]))
        worst["X"] = max(worst["X"], np.abs(o["X"][b] - r["X"]).max())
        worst["U"] = max(worst["U"], np.abs(o["U"][b] - r["U"]).max())
    frac = same_path / X0.shape[0]
    print(f"{spec.name} constrained={constrained}: same verdict and iteration counts {same_path}/{X0.shape[0]}, "
          f"worst differences among those {worst}")
    assert frac == 1.0, frac
    assert worst == dict(cost=0.0, X=0.0, U=0.0, viol=0.0), worst
    return frac, worst


def test_reference_build_reproduces_its_own_golden_values(ref):
    # test/ilqr/unicycle_ilqr_test.cpp: 9 iterations; test/augmented_lagrangian/auglag_test.cpp:346-350: 14 / 5 and the cost
    spec = P.unicycle_problem(P.K_TURN90)
    r = ref_solve(ref, "altro_ref_unicycle", 0, False, spec.x0, 3, 2, 100)
    assert (r["status"], r["total"]) == (0, 9) and abs(r["initial_cost"] - 259.27636137767087) < 1e-5
    r = ref_solve(ref, "altro_ref_unicycle", 0, True, spec.x0, 3, 2, 100, [1e-6, -1, -1, -1, -1, -1])
    assert (r["status"], r["outer"], r["total"]) == (0, 5, 14)
    assert abs(r["cost"] - 0.03893465058924039) / 0.03893465058924039 < 1e-11 and r["viol"] < 1e-6
    # perf/benchmark_unicycle.cpp: the three-obstacle solve takes 50 iterations
    r = ref_solve(ref, "altro_ref_unicycle", 1, True, spec.x0, 3, 2, 100, [-1, 10.0, -1, -1, -1, -1])
    assert (r["status"], r["total"]) == (0, 50)


@pytest.mark.parametrize("scenario,constrained", [(P.K_TURN90, False), (P.K_TURN90, True), (P.K_THREE_OBSTACLES, True)])
def test_oracle_follows_the_reference_build_unicycle(ref, scenario, constrained):
    spec = P.unicycle_problem(scenario)
    X0 = P.perturbed_initial_states(spec, 96, P.UNICYCLE_X0_SCALE)
    compare(ref, spec, "altro_ref_unicycle", scenario, constrained, X0)


def test_oracle_follows_the_reference_build_with_other_options(ref):
    # tighter tolerance, an explicit SetPenalty with initial_penalty = 0 (so that it is kept), iteration caps
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, 32, P.UNICYCLE_X0_SCALE, first=1000)
    o = ob.default_options()
    o.constraint_tolerance = 1e-6
    o.initial_penalty = 0.0
    o.max_iterations_total = 120
    o.max_iterations_inner = 40
    for b in range(X0.shape[0]):
        s = ob.OracleSolver(spec, use_constraints=True, options=o)
        s.set_initial_state(X0[b])
        s.set_penalty(25.0)
        s.solve_al()
        r = ref_solve(ref, "altro_ref_unicycle", 1, True, X0[b], 3, 2, 100, [1e-6, 25.0, 0.0, 120, 40, -1])
        st = s.status()
        assert (st["status"], st["iterations_outer"], st["iterations_total"]) == (r["status"], r["outer"], r["total"]), (b, st, r)
        X, U = s.trajectory()
        assert np.array_equal(X, r["X"]) and np.array_equal(U, r["U"])
        assert s.max_penalty() == r["max_penalty"]


@pytest.mark.parametrize("constrained", [False, True])
def test_oracle_follows_the_reference_build_triple_integrator(ref, constrained):
    spec = P.triple_integrator_problem(dof=2, N=50, add_constraints=constrained)
    X0 = P.perturbed_initial_states(spec, 64, P.TRIPLE_INTEGRATOR_X0_SCALE)
    compare(ref, spec, "altro_ref_triple_integrator", 50, constrained, X0)


@pytest.mark.parametrize("name,entry,first,constrained", [
    ("c1_unicycle_turn90_ilqr", "altro_ref_unicycle", 0, False),
    ("c2_unicycle_three_obstacles_al", "altro_ref_unicycle", 1, True),
    ("c3_triple_integrator_al", "altro_ref_triple_integrator", 50, True),
])
def test_committed_golden_vectors_are_what_the_reference_build_produces(ref, name, entry, first, constrained):
    """tests/golden/*.npz were written by the oracle (tests/golden/make_golden.py); the GPU box checks the CUDA path
    against them.  Here the reference's own code is asked for the same instances: the files hold its outputs."""
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    n, m = g["X"].shape[2], g["U"].shape[2]
    N = g["U"].shape[1]
    for b in range(g["X0"].shape[0]):
        r = ref_solve(ref, entry, first, constrained, g["X0"][b], n, m, N)
        assert r["status"] == int(g["status"][b]), b
        assert (r["outer"], r["total"]) == (int(g["iters"][b, 1]), int(g["iters"][b, 2])), b
        # bit-equal on the host that generated them; libm picks CPU-specific sin/cos kernels, so allow last-bit drift elsewhere
        for key, got in (("X", r["X"]), ("U", r["U"]), ("cost", r["cost"])):
            want = g[key][b]
            assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max()), (b, key)


# ---------------------------------------------------------------------------------------------------------------
# Any ProblemSpec on the reference's solver: oracle/ref_shim/ref_driver.cpp replays the builder calls of
# altro_cpp_b200/problems.py (the same calls build the oracle's and the device's problem) into the reference's own
# Problem / QuadraticCost / constraint classes.  The cart-pole (C4) and the discrete linear system (C5) are not in
# the reference: there they are user functors against its ABCs, with the closed forms of the oracle.
# ---------------------------------------------------------------------------------------------------------------
def ref_generic(lib, spec, constrained, x0, U0=None, options=None):
    handle = spec.build(lib, "altro_refb_")
    n, m, N = spec.n, spec.m, spec.N
    X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
    U0 = np.ascontiguousarray(spec.initial_controls() if U0 is None else U0, dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    opt = None if options is None else np.ascontiguousarray(options, dtype=np.float64)
    lib.altro_refb_solve.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 7
    got = lib.altro_refb_solve(handle, int(constrained), _ptr(x0), _ptr(U0), _ptr(opt), _ptr(X), _ptr(U), _ptr(sc), _ptr(it))
    lib.altro_refb_problem_destroy.argtypes = [ctypes.c_void_p]
    lib.altro_refb_problem_destroy(handle)
    assert got == N
    return dict(X=X, U=U, cost=sc[0], viol=sc[1], max_penalty=sc[2], status=int(it[0]), inner=int(it[1]),
                outer=int(it[2]), total=int(it[3]))


def compare_generic(lib, spec, constrained, X0, expect_statuses=None):
    o = ob.solve_batch(spec, X0, use_al=constrained, nthreads=8, want_gains=False)
    statuses = set()
    for b in range(X0.shape[0]):
        r = ref_generic(lib, spec, constrained, X0[b])
        mine = (int(o["status"][b]), int(o["iters"][b, 0]), int(o["iters"][b, 1]), int(o["iters"][b, 2]))
        assert mine == (r["status"], r["inner"], r["outer"], r["total"]), (b, mine, r)
        assert o["cost"][b] == r["cost"] and o["viol"][b] == r["viol"], (b, o["cost"][b], r["cost"])
        assert np.array_equal(o["X"][b], r["X"]) and np.array_equal(o["U"][b], r["U"]), b
        statuses.add(r["status"])
    print(f"{spec.name}: {X0.shape[0]} instances bit-identical; verdicts seen {sorted(statuses)}, "
          f"iterations {int(o['iters'][:, 2].min())}..{int(o['iters'][:, 2].max())}")
    if expect_statuses is not None:
        assert statuses == set(expect_statuses), statuses


def test_generic_builder_matches_the_reference_problem_factories(ref):
    # the replayed ProblemSpec and the reference's own examples/problems/unicycle.cpp give the same solve
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, 6, P.UNICYCLE_X0_SCALE)
    for b in range(6):
        a = ref_generic(ref, spec, True, X0[b])
        f = ref_solve(ref, "altro_ref_unicycle", 1, True, X0[b], 3, 2, 100)
        assert (a["status"], a["outer"], a["total"], a["cost"]) == (f["status"], f["outer"], f["total"], f["cost"])
        assert np.array_equal(a["X"], f["X"]) and np.array_equal(a["U"], f["U"])


def test_oracle_follows_the_reference_solver_on_the_cartpole_c4(ref):
    # BASELINE config C4: 100 inner iterations, kMaxInnerIterations — line-search failures and regularisation at work
    spec = P.cartpole_problem()
    compare_generic(ref, spec, True, P.perturbed_initial_states(spec, 10, P.CARTPOLE_X0_SCALE))


@pytest.mark.parametrize("literal", [False, True])
def test_oracle_follows_the_reference_solver_on_the_random_lqr_c5(ref, literal):
    # BASELINE config C5 (n = 32, m = 8): the m = 8 Cholesky; `literal` is the ill-conditioned variant whose first
    # backward pass fails its LLT at zero regularisation (restarts, Increase/DecreaseRegularization)
    spec = P.random_lqr_problem(literal=literal)
    compare_generic(ref, spec, True, P.normal_initial_states(spec, 6 if literal else 10))


def test_oracle_follows_the_reference_solver_on_a_stretched_time_grid(ref):
    # per-knot steps (Trajectory::SetStep / SetTime): h_k grows along the horizon
    spec = P.unicycle_problem(P.K_TURN90)
    N = spec.N
    h = (np.float32(0.02) + np.float32(0.0002) * np.arange(N + 1, dtype=np.float32)).astype(np.float32)
    h[N] = 0.0
    t = np.zeros(N + 1, dtype=np.float32)
    for k in range(N):
        t[k + 1] = np.float32(t[k] + h[k])
    spec.set_steps(t, h)
    compare_generic(ref, spec, True, P.perturbed_initial_states(spec, 8, P.UNICYCLE_X0_SCALE))


def test_oracle_follows_the_reference_solver_on_a_one_dof_triple_integrator(ref):
    # n = 3, m = 1: the run-time sized instantiation of the reference's templates
    spec = P.triple_integrator_problem(dof=1, N=30, add_constraints=True)
    compare_generic(ref, spec, True, P.perturbed_initial_states(spec, 8, (0.5,) * 3))


def ref_generic_ex(lib, spec, constrained, x0, options=None, history_cap=400):
    handle = spec.build(lib, "altro_refb_")
    n, m, N = spec.n, spec.m, spec.N
    X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
    K = np.zeros((N, n, m)); d = np.zeros((N, m)); hist = np.zeros((history_cap, 8)); rows = ctypes.c_int(0)
    U0 = np.ascontiguousarray(spec.initial_controls(), dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    opt = None if options is None else np.ascontiguousarray(options, dtype=np.float64)
    lib.altro_refb_solve_ex.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 10 + [ctypes.c_int, ctypes.c_void_p]
    lib.altro_refb_solve_ex(handle, int(constrained), _ptr(x0), _ptr(U0), _ptr(opt), _ptr(X), _ptr(U), _ptr(sc), _ptr(it),
                            _ptr(K), _ptr(d), _ptr(hist), history_cap, ctypes.byref(rows))
    lib.altro_refb_problem_destroy.argtypes = [ctypes.c_void_p]
    lib.altro_refb_problem_destroy(handle)
    return dict(X=X, U=U, K=np.ascontiguousarray(K.transpose(0, 2, 1)), d=d, history=hist[:rows.value], cost=sc[0])


@pytest.mark.parametrize("config", ["unicycle-3obs", "triple-integrator", "unicycle-ilqr"])
def test_gains_and_solver_stats_vectors_match_the_reference_build(ref, config):
    """The feedback / feed-forward gains a solve leaves behind and the per-iteration SolverStats vectors (cost, alpha,
    improvement_ratio, gradient, cost_decrease, regularization, violations, max_penalty; solver_stats.hpp:54-61, with
    the carry-forward rows of Log / NewIteration) — the device is tested against the oracle's on both counts."""
    constrained = config != "unicycle-ilqr"
    if config == "triple-integrator":
        spec = P.triple_integrator_problem(dof=2, N=50, add_constraints=True)
        X0 = P.perturbed_initial_states(spec, 6, P.TRIPLE_INTEGRATOR_X0_SCALE)
    else:
        spec = P.unicycle_problem(P.K_THREE_OBSTACLES if constrained else P.K_TURN90)
        X0 = P.perturbed_initial_states(spec, 6, P.UNICYCLE_X0_SCALE)
    names = ("cost", "alpha", "z", "gradient", "cost_decrease", "regularization", "violations", "max_penalty")
    for b in range(X0.shape[0]):
        r = ref_generic_ex(ref, spec, constrained, X0[b])
        s = ob.OracleSolver(spec, use_constraints=constrained)
        s.set_initial_state(X0[b])
        (s.solve_al if constrained else s.solve_ilqr)()
        K, d = s.gains()
        assert np.array_equal(K, r["K"]) and np.array_equal(d, r["d"]), b
        for col, name in enumerate(names):
            mine = s.stat(name)
            assert mine.shape[0] == r["history"].shape[0], (b, name, mine.shape, r["history"].shape)
            assert np.array_equal(mine, r["history"][:, col]), (b, name)


def test_warm_started_resolve_matches_the_reference_build(ref):
    """MPC-style re-solve (docs/Overview.dox:49-54 there; solver_options.hpp:47-48): a second and a third Solve() on the
    same solver, from the previous solution, with reset_duals = false and initial_penalty = 0 — multipliers and
    penalties carry over (al_solver.hpp:292-297).  The device's warm start is tested against the oracle's."""
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, 8, P.UNICYCLE_X0_SCALE)
    n, m, N = spec.n, spec.m, spec.N
    ref.altro_refb_solve_warm.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 4
    ref.altro_refb_problem_destroy.argtypes = [ctypes.c_void_p]
    for b in range(X0.shape[0]):
        for warm in (1, 2):
            handle = spec.build(ref, "altro_refb_")
            X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
            U0 = np.ascontiguousarray(spec.initial_controls(), dtype=np.float64)
            x0 = np.ascontiguousarray(X0[b], dtype=np.float64)
            ref.altro_refb_solve_warm(handle, _ptr(x0), _ptr(U0), None, warm, _ptr(X), _ptr(U), _ptr(sc), _ptr(it))
            ref.altro_refb_problem_destroy(handle)
            s = ob.OracleSolver(spec, use_constraints=True)
            s.set_initial_state(X0[b])
            s.solve_al()
            o = ob.default_options()
            o.reset_duals = 0
            o.initial_penalty = 0.0
            s.set_options(o)
            for _ in range(warm):
                s.solve_al()
            st = s.status()
            assert (st["status"], st["iterations_inner"], st["iterations_outer"], st["iterations_total"]) == tuple(int(v) for v in it), (b, warm, st, it)
            Xo, Uo = s.trajectory()
            assert np.array_equal(Xo, X) and np.array_equal(Uo, U), (b, warm)
            assert s.max_penalty() == sc[2]


# ---------------------------------------------------------------------------------------------------------------
# Fuzz: random problems (horizon, weights, goal, control bounds with infinite sides, 0-3 circular obstacles, uniform or
# jittered time grids, initial controls), random termination options; every one must come out of the oracle exactly
# as it comes out of the reference's own solver — whatever the verdict.
# ---------------------------------------------------------------------------------------------------------------
def random_unicycle_problem(rng):
    import math
    N = int(rng.integers(8, 60))
    n, m = 3, 2
    spec = P.ProblemSpec(n, m, N, name=f"fuzz-unicycle-N{N}")
    spec.set_model(P.MODEL_UNICYCLE, [])
    h = np.float32(np.float32(rng.uniform(1.0, 4.0)) / np.float32(N))
    if rng.random() < 0.5:
        spec.set_uniform_step(h)
    else:
        hs = (h * (1 + 0.3 * rng.uniform(-1, 1, N + 1))).astype(np.float32)
        hs[N] = 0
        t = np.zeros(N + 1, dtype=np.float32)
        for k in range(N):
            t[k + 1] = np.float32(t[k] + hs[k])
        spec.set_steps(t, hs)
    Q = np.diag(rng.uniform(1e-3, 1.0, n)); R = np.diag(rng.uniform(1e-3, 1.0, m)); Qf = np.diag(rng.uniform(1.0, 500.0, n))
    xf = np.array([rng.uniform(0.5, 2.5), rng.uniform(0.5, 2.5), rng.uniform(-math.pi, math.pi)])
    uref = rng.uniform(-0.2, 0.2, m)
    spec.set_cost(0, N, *P.lqr_cost(Q, R, xf, uref))
    spec.set_cost(N, N + 1, *P.lqr_cost(Qf, R * 0, xf, uref))
    obstacles = int(rng.integers(0, 4))
    cx, cy, cr = rng.uniform(0.2, 2.2, obstacles), rng.uniform(0.2, 2.2, obstacles), rng.uniform(0.05, 0.35, obstacles)
    lb = np.array([rng.uniform(-1.0, 0.0), -rng.uniform(0.5, 3.0)])
    ub = np.array([rng.uniform(0.8, 3.0), rng.uniform(0.5, 3.0)])
    if rng.random() < 0.3:
        lb[0] = -np.inf
    if rng.random() < 0.2:
        ub[1] = np.inf
    for k in range(N):
        if obstacles and k >= 1:
            spec.add_circles(k, cx, cy, cr)
        spec.add_control_bound(k, lb, ub)
    if rng.random() < 0.8:
        spec.add_goal(N, xf)
    spec.set_initial_state(np.zeros(n))
    spec.u0 = rng.uniform(-0.3, 0.5, m)
    spec.xf = xf
    return spec


def test_fuzzed_problems_come_out_of_the_oracle_as_out_of_the_reference_solver(ref):
    rng = np.random.default_rng(20251017)
    verdicts = {}
    for trial in range(160):
        spec = random_unicycle_problem(rng)
        x0 = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.6, 0.6)])
        o = ob.default_options()
        opts = [-1.0] * 6
        if rng.random() < 0.5:
            o.constraint_tolerance = opts[0] = float(10 ** rng.uniform(-7, -3))
        if rng.random() < 0.3:
            o.max_iterations_total = int(rng.integers(5, 60)); opts[3] = o.max_iterations_total
        if rng.random() < 0.3:
            o.max_iterations_inner = int(rng.integers(3, 30)); opts[4] = o.max_iterations_inner
        if rng.random() < 0.3:
            o.max_iterations_outer = int(rng.integers(1, 6)); opts[5] = o.max_iterations_outer
        s = ob.OracleSolver(spec, use_constraints=True, options=o)
        s.set_initial_state(x0)
        s.solve_al()
        st = s.status()
        X, U = s.trajectory()
        r = ref_generic(ref, spec, True, x0, options=opts)
        mine = (st["status"], st["iterations_inner"], st["iterations_outer"], st["iterations_total"])
        assert mine == (r["status"], r["inner"], r["outer"], r["total"]), (trial, spec.name, st, r)
        assert np.array_equal(X, r["X"]) and np.array_equal(U, r["U"]) and s.cost() == r["cost"], (trial, spec.name)
        verdicts[r["status"]] = verdicts.get(r["status"], 0) + 1
    print("verdicts seen (SolverStatus -> count):", dict(sorted(verdicts.items())))
    # solved, hit the total / outer / inner iteration limits, ran into the maximum penalty
    assert set(verdicts) >= {0, 5, 6, 7, 8}, verdicts


def _set_extra(lib, extra):
    """state_max, control_max, bp_reg_max, bp_reg_fail_threshold, line_search_max_iterations, cost_tolerance,
    gradient_tolerance, bp_reg_initial for all following reference solves (None: back to the defaults)"""
    if extra is None:
        lib.altro_ref_set_extra_options(None)
    else:
        e = np.ascontiguousarray(extra, dtype=np.float64)
        lib.altro_ref_set_extra_options(e.ctypes.data_as(ctypes.c_void_p))


def test_fuzzed_limits_and_line_search_options(ref):
    """state / control limits in the forward pass (kStateLimit, kControlLimit), short line searches, other convergence
    tolerances: the rarely taken branches of ilqr.hpp:468-558."""
    rng = np.random.default_rng(7)
    verdicts = {}
    try:
        for trial in range(160):
            spec = random_unicycle_problem(rng)
            x0 = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.6, 0.6)])
            o = ob.default_options()
            extra = [-1.0] * 8
            mode = trial % 4
            if mode == 0:
                o.state_max = extra[0] = float(rng.uniform(1.0, 3.5))
            elif mode == 1:
                o.control_max = extra[1] = float(rng.uniform(0.5, 3.0))
            elif mode == 2:
                o.line_search_max_iterations = int(rng.integers(1, 5)); extra[4] = o.line_search_max_iterations
            else:
                o.cost_tolerance = extra[5] = float(10 ** rng.uniform(-8, -2))
                o.gradient_tolerance = extra[6] = float(10 ** rng.uniform(-5, -1))
            _set_extra(ref, extra)
            s = ob.OracleSolver(spec, use_constraints=True, options=o)
            s.set_initial_state(x0)
            s.solve_al()
            st = s.status()
            X, U = s.trajectory()
            r = ref_generic(ref, spec, True, x0)
            mine = (st["status"], st["iterations_inner"], st["iterations_outer"], st["iterations_total"])
            assert mine == (r["status"], r["inner"], r["outer"], r["total"]), (trial, mode, st, r)
            assert np.array_equal(X, r["X"]) and np.array_equal(U, r["U"]), (trial, mode)
            verdicts[r["status"]] = verdicts.get(r["status"], 0) + 1
    finally:
        _set_extra(ref, None)
    print("verdicts seen (SolverStatus -> count):", dict(sorted(verdicts.items())))
    assert set(verdicts) >= {0, 2, 3, 7}, verdicts


def test_capped_regularisation_on_the_ill_conditioned_lqr(ref):
    """C5 'literal': the first backward pass fails its Cholesky; with bp_reg_max capped the regularisation saturates
    (ilqr.hpp:401-442, :770-786) and the solve takes another route — the same one in the oracle."""
    spec = P.random_lqr_problem(literal=True)
    X0 = P.normal_initial_states(spec, 3)
    try:
        for reg_max, threshold in [(1e-10, 1), (1e-9, 2), (1e-3, 5)]:
            o = ob.default_options()
            o.bp_reg_max = reg_max
            o.bp_reg_fail_threshold = threshold
            extra = [-1.0] * 8
            extra[2], extra[3] = reg_max, threshold
            _set_extra(ref, extra)
            for b in range(X0.shape[0]):
                s = ob.OracleSolver(spec, use_constraints=True, options=o)
                s.set_initial_state(X0[b])
                s.solve_al()
                st = s.status()
                X, U = s.trajectory()
                r = ref_generic(ref, spec, True, X0[b])
                mine = (st["status"], st["iterations_inner"], st["iterations_outer"], st["iterations_total"])
                assert mine == (r["status"], r["inner"], r["outer"], r["total"]), (reg_max, threshold, b, st, r)
                assert np.array_equal(X, r["X"]) and np.array_equal(U, r["U"]), (reg_max, threshold, b)
    finally:
        _set_extra(ref, None)


def test_fuzzed_triple_integrators_and_cartpoles(ref):
    """random triple integrators (one and two degrees of freedom: n = 3 and n = 6) with control bounds and a goal, and
    cart-poles over random horizons, default options"""
    rng = np.random.default_rng(11)
    verdicts = {}
    for trial in range(60):
        if trial % 3 < 2:
            dof = int(rng.integers(1, 3))
            n, m, N = 3 * dof, dof, int(rng.integers(5, 60))
            spec = P.ProblemSpec(n, m, N, name=f"fuzz-triple-integrator-dof{dof}-N{N}")
            spec.set_model(P.MODEL_TRIPLE_INTEGRATOR, [])
            spec.set_uniform_step(np.float32(rng.uniform(0.02, 0.2)))
            Q = np.diag(rng.uniform(0.01, 2.0, n)); R = np.diag(rng.uniform(1e-4, 0.1, m)); Qf = np.diag(rng.uniform(10, 1e5, n))
            xf = np.concatenate([rng.uniform(-2, 2, dof), np.zeros(2 * dof)])
            spec.set_cost(0, N, *P.lqr_cost(Q, R, xf, np.zeros(m)))
            spec.set_cost(N, N + 1, *P.lqr_cost(Qf, R * 0, xf, np.zeros(m)))
            ub = rng.uniform(1, 200, m)
            for k in range(N):
                spec.add_control_bound(k, -ub, ub)
            if rng.random() < 0.8:
                spec.add_goal(N, xf)
            spec.set_initial_state(np.zeros(n))
            spec.u0 = np.zeros(m)
            spec.xf = xf
            x0 = rng.uniform(-1.5, 1.5, n)
        else:
            spec = P.cartpole_problem(N=int(rng.integers(20, 120)))
            x0 = rng.uniform(-0.2, 0.2, 4)
        s = ob.OracleSolver(spec, use_constraints=True)
        s.set_initial_state(x0)
        s.solve_al()
        st = s.status()
        X, U = s.trajectory()
        r = ref_generic(ref, spec, True, x0)
        mine = (st["status"], st["iterations_inner"], st["iterations_outer"], st["iterations_total"])
        assert mine == (r["status"], r["inner"], r["outer"], r["total"]), (trial, spec.name, st, r)
        assert np.array_equal(X, r["X"]) and np.array_equal(U, r["U"]), (trial, spec.name)
        verdicts[r["status"]] = verdicts.get(r["status"], 0) + 1
    print("verdicts seen (SolverStatus -> count):", dict(sorted(verdicts.items())))
