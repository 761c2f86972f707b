"""Reproducibility of the CUDA path at BASELINE.json's batch sizes (reference analogue: the re-solve
determinism check of test/augmented_lagrangian/auglag_test.cpp:353-380 and the nthreads-equivalence
tests of test/ilqr/ilqr_class_test.cpp:130-160).

Two FRESH solvers (other device addresses, no history) on identical inputs, driven like bench.py
drives them — device-resident inputs on a non-blocking stream — must agree bit for bit on every
output: status, iteration counters, cost, violation, X, U, K, d.  So must a solver with the opt-in
stall skip.  Compared as raw bits (NaN-safe).
"""
import numpy as np
import pytest

from altro_cpp_b200 import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (the product path has no CPU fallback)")
    import altro_cpp_b200 as pkg
    pkg.set_default_engine(None)
    return pkg


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int64) if a.dtype == np.float64 else a


def _solve_like_bench(gpu, spec, X0, options=None, solves=2):
    import torch
    B = X0.shape[0]
    dev = torch.device("cuda", 0)
    x0_dev = torch.from_numpy(X0).to(dev)
    stream = torch.cuda.Stream(device=dev)  # non-blocking, like bench.py
    s = gpu.BatchSolver(spec, B, options=options)
    with torch.cuda.stream(stream):
        for _ in range(solves):
            s.set_inputs_dev(x0_dev.data_ptr(), 0, spec.u0, stream=stream)
            s.solve_al(stream=stream)
        torch.cuda.synchronize(dev)
    r = s.results()
    X, U = s.trajectory()
    K, d = s.gains()
    return dict(status=r["status"], iters=r["iters"], cost=r["cost"], viol=r["viol"], X=X, U=U, K=K, d=d)


def _assert_identical(a, b, B, what):
    bad = {k: int((_bits(a[k]) != _bits(b[k])).reshape(B, -1).any(axis=1).sum()) for k in a}
    assert not any(bad.values()), f"{what}: instances differing per field {bad}"


@pytest.mark.parametrize("case", ["c2", "c3", "c4"])
def test_fresh_solvers_are_bit_identical_at_baseline_batch(gpu, case):
    if case == "c2":
        spec, scale, B = P.unicycle_problem(P.K_THREE_OBSTACLES), P.UNICYCLE_X0_SCALE, 16384
    elif case == "c3":
        spec, scale, B = P.triple_integrator_problem(dof=2, N=50, add_constraints=True), P.TRIPLE_INTEGRATOR_X0_SCALE, 8192
    else:
        spec, scale, B = P.cartpole_problem(N=200), P.CARTPOLE_X0_SCALE, 32768
    X0 = P.perturbed_initial_states(spec, B, scale)
    a = _solve_like_bench(gpu, spec, X0, solves=1)   # first solve of a fresh solver: lazily allocated scratch
    b = _solve_like_bench(gpu, spec, X0, solves=3)   # third solve of another one
    _assert_identical(a, b, B, "two fresh solvers")
    o = gpu.default_options()
    o.skip_repeated_iterations = 1
    c = _solve_like_bench(gpu, spec, X0, options=o, solves=1)
    _assert_identical(a, c, B, "skip_repeated_iterations")


def test_backward_pass_after_solve_needs_update_expansions(gpu):
    """The records left in EXP by a whole solve are per-slot scratch, not the expansion of Z_:
    BackwardPass must ask for UpdateExpansions instead of running on them (and must not touch
    cost-to-go arrays that were never allocated)."""
    spec = P.unicycle_problem(P.K_TURN90)
    X0 = P.perturbed_initial_states(spec, 16, P.UNICYCLE_X0_SCALE)
    s = gpu.BatchSolver(spec, 16, use_constraints=False)
    s.set_inputs(X0)
    s.solve_ilqr()
    with pytest.raises(gpu.SolverError, match="UpdateExpansions"):
        s.backward_pass()
    s.update_expansions()
    s.backward_pass()
    P0, p0 = s.ctg(0)
    assert np.all(np.isfinite(P0)) and np.all(np.isfinite(p0))


def test_dense_and_diagonal_cost_paths_agree_bitwise(gpu, monkeypatch):
    """The diagonal-cost fast path drops exact-zero products only: same bits as the dense path."""
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, 96, P.UNICYCLE_X0_SCALE)
    a = _solve_like_bench(gpu, spec, X0, solves=1)
    monkeypatch.setenv("ALTRO_B200_DENSE_COST", "1")
    b = _solve_like_bench(gpu, spec, X0, solves=1)
    _assert_identical(a, b, 96, "dense vs diagonal cost evaluation")



def test_sharded_solver_is_bit_identical_to_single(gpu):
    """altro_b200_multi_*: the batch cut into contiguous slices, one solver + host thread per entry of
    `devices` (here: three slices on the one GPU a test box has; on an 8-GPU node the same code path with
    devices = range(8)).  An instance's result does not depend on its slice: bit-identical to one solver."""
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    B = 1000  # ragged: 334 + 333 + 333
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    one = gpu.BatchSolver(spec, B)
    a = one.solve_al_host(X0)
    sharded = gpu.MultiBatchSolver(spec, B, devices=[0, 0, 0])
    b = sharded.solve_al_host(X0)
    for k in ("status", "iters", "cost", "viol", "X", "U"):
        assert np.array_equal(_bits(a[k]), _bits(b[k])), k
    t = sharded.timings()
    assert t["solve_ms"].shape == (3,) and np.all(t["solve_ms"] > 0)


@pytest.mark.parametrize("case", ["c2", "c3", "c4"])
def test_results_do_not_depend_on_other_kernels_running(gpu, case):
    """Round-1/2 defect: the TMA ring of the backward-pass kernel refilled a slot right after reading it, with
    nothing ordering the async-proxy write after the generic-proxy loads; when unrelated kernels kept the SM's
    load/store pipeline busy, a refill that hit L2 overtook the last loads of a slot and one tile read rows of the
    wrong knot (18 of 20 solves differed under this load; tools/gpu_flaky*.py).  Repeated solves while copies and
    GEMMs run on another stream must be bit-identical to a solve on the idle GPU."""
    import threading
    import torch
    if case == "c2":    # k_backward_mat<Unicycle>
        spec, scale, B, reps = P.unicycle_problem(P.K_THREE_OBSTACLES), P.UNICYCLE_X0_SCALE, 1000, 12
    elif case == "c3":  # k_backward_coop (two lanes per instance, same ring)
        spec, scale, B, reps = P.triple_integrator_problem(dof=2, N=50, add_constraints=True), P.TRIPLE_INTEGRATOR_X0_SCALE, 4096, 40
    else:               # k_backward_mat of a run-time compiled model
        spec, scale, B, reps = P.cartpole_problem(N=200), P.CARTPOLE_X0_SCALE, 512, 6
    X0 = P.perturbed_initial_states(spec, B, scale)
    s = gpu.BatchSolver(spec, B)
    ref = s.solve_al_host(X0)
    dev = torch.device("cuda", 0)
    stop = []

    def load():
        s2 = torch.cuda.Stream(device=dev)
        a = torch.empty(1 << 26, dtype=torch.float64, device=dev)
        b = torch.empty_like(a)
        m1 = torch.randn(4096, 4096, device=dev, dtype=torch.float32)
        with torch.cuda.stream(s2):
            while not stop:
                for _ in range(4):
                    b.copy_(a)
                    m1 = (m1 @ m1).clamp_(-1, 1)
                s2.synchronize()

    t = threading.Thread(target=load, daemon=True)
    t.start()
    try:
        for rep in range(reps):
            out = s.solve_al_host(X0)
            bad = {k: int((_bits(out[k]) != _bits(ref[k])).reshape(B, -1).any(axis=1).sum())
                   for k in ("status", "iters", "cost", "viol", "X", "U")}
            assert not any(bad.values()), f"solve {rep} under load: instances differing per field {bad}"
    finally:
        stop.append(1)
        t.join(timeout=20)
