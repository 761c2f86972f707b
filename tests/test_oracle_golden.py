"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests.

Every expected value below is copied from a test in /root/reference (file:line in
the comment) together with the tolerance that test uses.  Most values were generated
by Altro.jl with a double time step; the C++ reference (and this oracle) computes
with ``float h``, which is why they agree to ~1e-8 relative rather
than to the last bit — the same margin the reference's tests allow.
"""
import numpy as np
import pytest

from altro_cpp_b200 import problems as P

SOLVED, UNSOLVED = 0, 1


def isapprox(a, b, prec):
    """Eigen's isApprox: ||a-b|| <= prec * min(||a||, ||b||)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


# ---------------------------------------------------------------- unicycle kTurn90 (iLQR)
def make_unicycle(oracle, al=False, scenario=P.K_TURN90, **kw):
    s = oracle.OracleSolver(P.unicycle_problem(scenario, **kw), use_constraints=al)
    s.rollout()  # UnicycleProblem::MakeSolver ends with solver.Rollout()
    return s


def test_unicycle_initialization(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:32-37
    s = make_unicycle(oracle)
    assert abs(s.cost() - 259.27636137767087) < 1e-5


def test_unicycle_backward_pass(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:39-54
    s = make_unicycle(oracle)
    s.update_expansions()
    s.backward_pass()
    ctg_grad0 = [0.024904637422419617, -0.46496022574032614, -0.0573096310550007]
    d0 = [-2.565783457444465, 5.514158930898376]
    assert isapprox(s.ctg(0)[1], ctg_grad0, 1e-5)
    assert isapprox(s.gains()[1][0], d0, 1e-5)


def test_unicycle_forward_pass(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:56-65  (EXPECT_DOUBLE_EQ alpha)
    s = make_unicycle(oracle)
    s.update_expansions()
    s.backward_pass()
    J0 = s.cost()
    s.forward_pass()
    assert s.cost() < J0
    assert s.stat("alpha")[0] == 0.0625


def test_unicycle_two_steps(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:67-88
    s = make_unicycle(oracle)
    s.update_expansions(); s.backward_pass(); s.forward_pass()
    s.update_expansions(); s.backward_pass()
    ctg_grad0 = [-0.0015143873973949232, -0.07854630832127288, -0.017945283678268698]
    d0 = [0.21887571453613042, 1.3097976615154625]
    assert isapprox(s.ctg(0)[1], ctg_grad0, 1e-5)
    assert isapprox(s.gains()[1][0], d0, 1e-5)
    s.forward_pass()
    assert abs(s.cost() - 62.773696055304384) < 1e-5


def test_unicycle_full_solve(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:90-100
    s = make_unicycle(oracle)
    s.solve_ilqr()
    st = s.status()
    assert st["iterations_inner"] == 9
    assert st["status"] == SOLVED
    assert abs(s.cost() - 0.0387016567) < 1e-5
    assert s.stat("gradient")[-1] < 1e-2


def test_unicycle_auglag_forward_pass(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:102-113
    s = make_unicycle(oracle, al=True)
    s.update_expansions(); s.backward_pass()
    J0 = s.cost()
    s.forward_pass()
    assert s.cost() < J0
    assert s.stat("alpha")[0] == 0.0625


def test_unicycle_auglag_first_ilqr_solve(oracle):
    # test/ilqr/unicycle_ilqr_test.cpp:115-145 and test/augmented_lagrangian/auglag_test.cpp:224-247
    s = make_unicycle(oracle, al=True)
    s.solve_ilqr()
    J = s.cost()
    X, U = s.trajectory()
    viol = max(np.abs(U[:, 0]).max() - 1.5, np.abs(U[:, 1]).max() - 1.5)
    J_expected = 0.03893427133384412
    viol_expected = 0.00017691645708972636
    assert abs(J_expected - J) / J_expected < 1e-6
    assert abs(viol_expected - viol) / viol_expected < 1e-6
    assert abs(viol_expected - s.max_violation_stored()) / viol_expected < 1e-6
    assert s.status()["iterations_inner"] == 10


def test_auglag_two_solves(oracle):
    # test/augmented_lagrangian/auglag_test.cpp:249-301
    s = oracle.OracleSolver(P.unicycle_problem(P.K_TURN90), use_constraints=True)
    J0 = s.cost()
    viol0 = s.max_violation_stored()
    s.solve_ilqr()
    J = s.cost()
    viol = s.max_violation_stored()
    s.update_duals(); s.update_penalties()
    assert s.cost() > J
    assert viol < viol0
    assert J < J0
    s.solve_ilqr()
    s.cost()
    viol = s.max_violation_stored()
    assert abs(0.0000626 - viol) / 0.0000626 < 0.1
    assert s.status()["iterations_inner"] == 1


@pytest.mark.parametrize("repeat", [1, 2])
def test_auglag_full_solve(oracle, repeat):
    # test/augmented_lagrangian/auglag_test.cpp:326-351 (InitializeAndSolve) and :353-380 (SolveTwice)
    spec = P.unicycle_problem(P.K_TURN90)
    s = oracle.OracleSolver(spec, use_constraints=True)
    o = oracle.default_options()
    o.constraint_tolerance = 1e-6
    s.set_options(o)
    for _ in range(repeat):
        s.set_controls(spec.initial_controls())
        s.solve_al()
    st = s.status()
    viol = s.max_violation_stored()
    assert st["iterations_total"] == 14
    assert st["iterations_outer"] == 5
    assert st["status"] == SOLVED
    assert viol < 1e-6
    # EXPECT_DOUBLE_EQ in the reference (4 ULP on the reference's own build); the
    # restatement reproduces it to ~1e-12 relative.
    assert abs(s.cost() - 0.03893465058924039) / 0.03893465058924039 < 1e-11


# ---------------------------------------------------------------- unicycle kThreeObstacles
def test_three_obstacles_construction(oracle):
    # test/examples/example_unicycle_test.cpp:18-29
    s = make_unicycle(oracle, scenario=P.K_THREE_OBSTACLES)
    assert abs(s.cost() - 133.1151550141444) < 1e-6
    s = make_unicycle(oracle, al=True, scenario=P.K_THREE_OBSTACLES)
    assert abs(s.cost() - 141.9639680271223) < 1e-6


def test_three_obstacles_increase_penalty(oracle):
    # test/examples/example_unicycle_test.cpp:31-50
    s = make_unicycle(oracle, al=True, scenario=P.K_THREE_OBSTACLES)
    s.set_penalty(10.0)
    assert abs(s.cost() - 221.6032851439234) < 1e-6


def test_three_obstacles_solve_one_step(oracle):
    # test/examples/example_unicycle_test.cpp:52-67
    s = make_unicycle(oracle, al=True, scenario=P.K_THREE_OBSTACLES)
    s.set_penalty(10.0)
    s.solve_ilqr()
    s.update_duals(); s.update_penalties()
    lambdaN = np.array([0.43555910438329626, -0.5998598475208317, 0.0044282251970790935])
    assert isapprox(s.duals(100), -lambdaN, 1e-6)


def test_three_obstacles_solve_constrained(oracle):
    # test/examples/example_unicycle_test.cpp:69-89  (Q10: SetPenalty(10) is overridden by Init())
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    s = make_unicycle(oracle, al=True, scenario=P.K_THREE_OBSTACLES)
    s.set_penalty(10.0)
    s.solve_al()
    X, U = s.trajectory()
    cx = cy = np.array([0.25, 0.5, 0.75]) * 3.0
    for i in range(3):
        dist = np.sqrt((X[:, 0] - cx[i]) ** 2 + (X[:, 1] - cy[i]) ** 2) - 0.425
        assert dist.min() > -1e-3
    st = s.status()
    assert st["status"] == SOLVED
    s.cost()
    assert s.max_violation_stored() < 1e-4
    assert s.stat("cost_decrease")[-1] < 1e-4
    assert s.stat("gradient")[-1] < 1e-2
    # SURVEY.md 3.3 / BASELINE.md: 50 iLQR iterations in 5 AL iterations for the nominal instance
    assert (st["iterations_total"], st["iterations_outer"]) == (50, 5)


def test_three_obstacles_serial_equals_repeat(oracle):
    # analogue of test/examples/example_unicycle_test.cpp:91-166 (SolveParallel): two
    # independent solves of the same N=10 problem with initial_penalty=10 are identical.
    outs = []
    for _ in range(2):
        s = make_unicycle(oracle, al=True, scenario=P.K_THREE_OBSTACLES, N=10)
        o = oracle.default_options()
        o.initial_penalty = 10.0
        s.set_options(o)
        s.solve_al()
        outs.append((s.cost(), s.status()["iterations_total"], s.gains()[0]))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2])


# ---------------------------------------------------------------- triple integrator
def make_triple(oracle, al=False, rollout_zero=False):
    # fixture of test/ilqr/ilqr_test.cpp:21-108 (goal constraint only)
    spec = P.triple_integrator_problem(dof=2, N=10, goal_only=True)
    s = oracle.OracleSolver(spec, use_constraints=al)
    if rollout_zero:  # RolloutZeroControls :50-56
        s.set_states(np.tile(spec.x0, (11, 1)))
    return s, spec


def test_triple_dynamics_expansion(oracle):
    # test/ilqr/ilqr_test.cpp:147-180
    s, spec = make_triple(oracle, rollout_zero=True)
    s.update_expansions()
    A = np.array([[1, 0, 0.1, 0, 0.005, 0], [0, 1, 0, 0.1, 0, 0.005], [0, 0, 1, 0, 0.1, 0],
                  [0, 0, 0, 1, 0, 0.1], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1.0]])
    B = np.array([[1 / 6e3, 0], [0, 1 / 6e3], [5e-3, 0], [0, 5e-3], [0.1, 0], [0, 0.1]])
    for k in range(10):
        e = s.expansion(k)
        assert isapprox(e["A"], A, 1e-6)
        assert isapprox(e["B"], B, 1e-6)
    e0 = s.expansion(0)
    # CostExpansion test :129-145
    assert isapprox(e0["lxx"], np.eye(6), 1e-12)
    assert isapprox(e0["luu"], np.eye(2) * 0.001, 1e-12)
    assert isapprox(e0["lx"], spec.x0 - spec.xf, 1e-12)
    assert np.all(e0["lu"] == 0)
    eN = s.expansion(10)
    assert isapprox(eN["lxx"], np.eye(6) * 1e5, 1e-12)
    assert isapprox(eN["lx"], 1e5 * (spec.x0 - spec.xf), 1e-12)


def test_triple_backward_pass(oracle):
    # test/ilqr/ilqr_test.cpp:182-204
    s, _ = make_triple(oracle, rollout_zero=True)
    s.update_expansions(); s.backward_pass()
    ctg_grad0 = [-389.04658272629644, -778.0931654525915, -181.40881931288234, -362.81763862576514,
                 -9.704677110465038, -19.409354220930084]
    d0 = [127.9313782698078, 255.862756539616]
    assert isapprox(s.ctg(0)[1], ctg_grad0, 1e-4)
    assert isapprox(s.gains()[1][0], d0, 1e-4)


def test_triple_cost(oracle):
    # test/ilqr/ilqr_test.cpp:206-232  (EXPECT_DOUBLE_EQ)
    s, spec = make_triple(oracle)
    s.rollout()
    assert s.cost() == pytest.approx(100 + 1e6, rel=4e-16)
    s.set_states(np.tile(spec.xf, (11, 1)))
    s.set_controls(np.zeros((10, 2)))
    assert s.cost() == 0.0


def test_triple_forward_pass(oracle):
    # test/ilqr/ilqr_test.cpp:256-269
    s, _ = make_triple(oracle)
    s.rollout(); s.update_expansions(); s.backward_pass()
    J0 = s.cost()
    s.forward_pass()
    J = s.cost()
    assert J < J0
    assert abs(J - 1945.2329136) < 1e-3


K0_TRIPLE = np.array([[-63.9657, 0.0, -42.7673, 0.0, -11.5189, 0.0],
                      [0.0, -63.9657, 0.0, -42.7673, 0.0, -11.5189]])


@pytest.mark.parametrize("al", [False, True])
def test_triple_two_steps(oracle, al):
    # test/ilqr/ilqr_test.cpp:271-302 and :413-451 (AugLagTwoSteps)
    s, spec = make_triple(oracle, al=al)
    s.rollout()
    costs = [s.cost()]
    for _ in range(2):
        s.update_expansions(); s.backward_pass(); s.forward_pass()
        costs.append(s.cost())
    assert costs[1] - costs[2] < (1e-4 if al else 1e-10)
    K, d = s.gains()
    assert isapprox(K[0], K0_TRIPLE, 1e-4)
    assert np.linalg.norm(d[0]) < 1e-8
    if al:
        X, _ = s.trajectory()
        assert np.abs(X[10] - spec.xf).max() < 0.01


def test_triple_full_solve(oracle):
    # test/ilqr/ilqr_test.cpp:304-336 and test/examples/example_triple_integrator_test.cpp:16-37
    s, _ = make_triple(oracle)
    s.solve_ilqr()
    st = s.status()
    assert st["status"] == SOLVED
    assert st["iterations_inner"] == 2
    assert isapprox(s.gains()[0][0], K0_TRIPLE, 1e-3)
    assert s.stat("cost_decrease")[-1] < 1e-4
    assert s.stat("gradient")[-1] < 1e-2


def test_triple_auglag_cost_and_expansion(oracle):
    # test/ilqr/ilqr_test.cpp:338-383
    s, spec = make_triple(oracle, al=True)
    s.rollout()
    J_goal = np.sum((spec.x0 - spec.xf) ** 2) / 2
    assert s.cost() == pytest.approx(100 + 1e6 + J_goal, rel=4e-16)
    rho = 123.0
    s.set_penalty(rho)
    s.set_duals(10, np.full(6, 1.5))
    lam_bar = 1.5 - rho * (spec.x0 - spec.xf)
    s.update_expansions()
    e = s.expansion(10)
    assert isapprox(e["lxx"], np.eye(6) * (1e5 + rho), 1e-12)
    assert np.all(e["luu"] == 0)
    assert isapprox(e["lx"], 1e5 * (spec.x0 - spec.xf) - lam_bar, 1e-12)
    assert np.all(e["lu"] == 0)


def test_triple_auglag_backward_forward(oracle):
    # test/ilqr/ilqr_test.cpp:385-411
    s, _ = make_triple(oracle, al=True)
    s.rollout(); s.update_expansions(); s.backward_pass()
    ctg_grad0 = [-389.04659149197226, -778.0931829839444, -181.4088232963142, -362.8176465926284,
                 -9.704677322846152, -19.40935464569231]
    d0 = [127.93131544425611, 255.86263088851214]
    assert isapprox(s.ctg(0)[1], ctg_grad0, 1e-4)
    assert isapprox(s.gains()[1][0], d0, 1e-4)
    J0 = s.cost()
    s.forward_pass()
    J = s.cost()
    assert J < J0
    assert abs(J - 1945.232957449998) < 1e-3


def test_triple_constrained_example(oracle):
    # test/examples/example_triple_integrator_test.cpp:39-70
    spec = P.triple_integrator_problem(dof=2, N=10, add_constraints=True)
    s = oracle.OracleSolver(spec, use_constraints=True)
    s.solve_al()
    st = s.status()
    assert st["status"] == SOLVED
    assert s.stat("cost_decrease")[-1] < 1e-4
    assert s.stat("gradient")[-1] < 1e-2
    assert s.stat("violations")[-1] < 1e-4
    X, U = s.trajectory()
    assert np.abs(X[10] - spec.xf).max() < 1e-4
    ubnd = np.array([100.0, 200.0])
    assert isapprox(U[0], ubnd, 1e-12 ** 0.5 * 1e6 * 1e-6)  # Eigen isApprox default prec 1e-12
    assert isapprox(U[9], ubnd, 1e-6)


# ---------------------------------------------------------------- unit-level identities
def test_knot_point_identities(oracle):
    # test/ilqr/knot_point_functions_test.cpp:54-146: Q-expansion, gains, cost-to-go against
    # dense numpy re-derivations on the unicycle first backward pass.
    s = make_unicycle(oracle)
    s.update_expansions(); s.backward_pass()
    K, d = s.gains()
    for k in (99, 50, 0):
        e = s.expansion(k)
        Pn, pn = s.ctg(k + 1)
        A, B = e["A"], e["B"]
        Q = s.action_value(k)
        np.testing.assert_allclose(Q["Qxx"], e["lxx"] + A.T @ Pn @ A, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Q["Qxu"], e["lxu"] + A.T @ Pn @ B, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Q["Quu"], e["luu"] + B.T @ Pn @ B, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Q["Qx"], e["lx"] + A.T @ pn, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(Q["Qu"], e["lu"] + B.T @ pn, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(K[k], -np.linalg.solve(Q["Quu"], Q["Qxu"].T), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(d[k], -np.linalg.solve(Q["Quu"], Q["Qu"]), rtol=1e-9, atol=1e-12)
        Pk, pk = s.ctg(k)
        Pe = Q["Qxx"] + K[k].T @ Q["Quu"] @ K[k] + K[k].T @ Q["Qxu"].T + Q["Qxu"] @ K[k]
        pe = Q["Qx"] + K[k].T @ Q["Quu"] @ d[k] + K[k].T @ Q["Qu"] + Q["Qxu"] @ d[k]
        np.testing.assert_allclose(Pk, Pe, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(pk, pe, rtol=1e-10, atol=1e-13)


def test_rk4_jacobian_vs_finite_difference(oracle):
    # test/problem/*: discrete Jacobian checked against central differences (derivative_checker.hpp)
    s = make_unicycle(oracle)
    X, U = s.trajectory()
    s.update_expansions()
    k = 37
    e = s.expansion(k)
    eps = 1e-6
    Jfd = np.zeros((3, 5))
    for j in range(5):
        for sgn in (+1, -1):
            t = oracle.OracleSolver(P.unicycle_problem(P.K_TURN90, N=1), use_constraints=False)
            x0 = X[k].copy(); u0 = U[k].copy()
            if j < 3:
                x0[j] += sgn * eps
            else:
                u0[j - 3] += sgn * eps
            t.set_initial_state(x0)
            t.set_controls(u0[None, :])
            # N=1 problem has h = 3/1; rebuild with the same h as the N=100 problem instead
            Jfd[:, j] += sgn * _rk4_unicycle(x0, u0, np.float32(np.float32(3.0) / np.float32(100))) / (2 * eps)
    np.testing.assert_allclose(np.hstack([e["A"], e["B"]]), Jfd, rtol=1e-6, atol=1e-8)


def _rk4_unicycle(x, u, h):
    h = float(h)
    f = lambda x: np.array([u[0] * np.cos(x[2]), u[0] * np.sin(x[2]), u[1]])
    k1 = f(x); k2 = f(x + 0.5 * h * k1); k3 = f(x + 0.5 * h * k2); k4 = f(x + h * k3)
    return x + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6


def test_al_cost_gradient_hessian_vs_finite_difference(oracle):
    # test/augmented_lagrangian/auglag_test.cpp:49-93: AL cost expansion at knot 0 of the
    # unicycle problem with a violated velocity bound, rho = 1.1, lambda[2] = -rho*0.5.
    spec = P.unicycle_problem(P.K_TURN90, N=2)
    x = np.array([0.1, 0.2, np.pi / 3])
    u = np.array([1.5 + 0.5, 1.5 / 2])
    rho = 1.1

    def solver_at(z, al=True, dual=True):
        s = oracle.OracleSolver(spec, use_constraints=al)
        if al:
            s.set_penalty(rho)
            if dual:
                s.set_duals(0, [0, 0, -rho * 0.5, 0])
        X = np.zeros((3, 3)); X[0] = z[:3]
        Uc = np.zeros((2, 2)); Uc[0] = z[3:]
        s.set_states(X); s.set_controls(Uc)
        return s

    z = np.concatenate([x, u])
    s = solver_at(z)
    s.update_expansions()
    e = s.expansion(0)
    eps = 1e-5
    g = np.zeros(5)
    Hfd = np.zeros((5, 5))
    for j in range(5):
        dz = np.zeros(5); dz[j] = eps
        sp, sm = solver_at(z + dz), solver_at(z - dz)
        g[j] = (sp.cost() - sm.cost()) / (2 * eps)   # knots 1, 2 do not depend on z
        sp.update_expansions(); sm.update_expansions()
        ep, em = sp.expansion(0), sm.expansion(0)
        Hfd[:, j] = (np.concatenate([ep["lx"], ep["lu"]]) - np.concatenate([em["lx"], em["lu"]])) / (2 * eps)
    assert isapprox(e["lx"], g[:3], 1e-4)
    assert isapprox(e["lu"], g[3:], 1e-4)
    assert isapprox(e["lxx"], Hfd[:3, :3], 1e-4)
    assert np.linalg.norm(e["lxu"] - Hfd[:3, 3:]) < 1e-4
    assert isapprox(e["luu"], Hfd[3:, 3:], 1e-4)
    # value: J_al = J_cost + 0.5*rho*viol^2 at knot 0 (auglag_test.cpp:52-60), plus the goal
    # term 0.5*rho*|x_N - xf|^2 at the terminal knot (x_N = 0 here).
    J_al = solver_at(z, dual=False).cost()
    J_plain = solver_at(z, al=False).cost()
    expected = 0.5 * rho * 0.5 ** 2 + 0.5 * rho * float(spec.xf @ spec.xf)
    assert J_al - J_plain == pytest.approx(expected, rel=1e-12)
    # with the dual set: 0.5*((-2 rho v)^2 - (rho v)^2)/rho  (auglag_test.cpp:62-66)
    J_al2 = solver_at(z).cost()
    expected2 = 0.5 * ((2 * rho * 0.5) ** 2 - (rho * 0.5) ** 2) / rho + 0.5 * rho * float(spec.xf @ spec.xf)
    assert J_al2 - J_plain == pytest.approx(expected2, rel=1e-12)
