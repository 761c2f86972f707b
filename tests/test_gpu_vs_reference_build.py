"""The device path against the REFERENCE's own solver code: oracle/_ref/libaltro_ref.so is the reference's
altro/**/*.cpp and examples compiled where they lie in the development container (oracle/build_ref.py, on this repo's
Eigen stand-in) — the prebuilt library travels to the GPU box, /root/reference is not read here.  Same problem
definitions (the reference's examples/problems/*.hpp inside the library, altro_cpp_b200/problems.py on the device
side), same perturbed initial states, whole AL-iLQR solves through the C ABI.

Per instance: verdict and iteration counts must be the reference's (a handful of instances may take another discrete
path because the kernels contract multiply-adds: the fraction is asserted and printed); on those that follow it, cost,
violation, states and controls agree to rounding.
"""
import ctypes
import os

import numpy as np
import pytest

from altro_cpp_b200 import problems as P

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libaltro_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libaltro_ref.so was not built (needs /root/reference: python oracle/build_ref.py)")
    return ctypes.CDLL(LIB)


@pytest.fixture(scope="module", params=["phased", "fused"])
def gpu(request):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (the product path has no CPU fallback)")
    import altro_cpp_b200 as pkg
    pkg.set_default_engine(request.param)
    yield pkg
    pkg.set_default_engine(None)


_solved = {}  # (entry, first_arg, x0 bytes) -> result: both engines are compared with the same reference solves


def ref_solve(lib, entry, first_arg, constrained, x0, n, m, N):
    key = (entry, first_arg, constrained, np.ascontiguousarray(x0, dtype=np.float64).tobytes())
    if key not in _solved:
        _solved[key] = _ref_solve(lib, entry, first_arg, constrained, x0, n, m, N)
    return _solved[key]


def _ref_solve(lib, entry, first_arg, constrained, x0, n, m, N):
    X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    got = getattr(lib, entry)(ctypes.c_int(first_arg), ctypes.c_int(int(constrained)), ptr(x0), None, ptr(X), ptr(U),
                              ptr(sc), ptr(it))
    assert got == N
    return X, U, sc, it


@pytest.mark.parametrize("config", ["unicycle-3obs", "unicycle-turn90", "triple-integrator"])
def test_device_solves_follow_the_reference_build(gpu, ref, config):
    if config == "triple-integrator":
        spec = P.triple_integrator_problem(dof=2, N=50, add_constraints=True)
        X0 = P.perturbed_initial_states(spec, 40, P.TRIPLE_INTEGRATOR_X0_SCALE)
        entry, first = "altro_ref_triple_integrator", 50
    else:
        scenario = P.K_THREE_OBSTACLES if config == "unicycle-3obs" else P.K_TURN90
        spec = P.unicycle_problem(scenario)
        X0 = P.perturbed_initial_states(spec, 40, P.UNICYCLE_X0_SCALE)
        entry, first = "altro_ref_unicycle", scenario
    B = X0.shape[0]
    s = gpu.BatchSolver(spec, B)
    s.set_inputs(X0)
    s.solve_al()
    res = s.results()
    X, U = s.trajectory()
    same, worst = 0, dict(cost=0.0, viol=0.0, X=0.0, U=0.0)
    for b in range(B):
        Xr, Ur, sc, it = ref_solve(ref, entry, first, True, X0[b], spec.n, spec.m, spec.N)
        mine = (int(res["status"][b]), int(res["iters"][b, 1]), int(res["iters"][b, 2]))
        theirs = (int(it[0]), int(it[2]), int(it[3]))
        if mine != theirs:
            continue
        same += 1
        worst["cost"] = max(worst["cost"], abs(res["cost"][b] - sc[0]) / max(1.0, abs(sc[0])))
        worst["viol"] = max(worst["viol"], abs(res["viol"][b] - sc[1]))
        worst["X"] = max(worst["X"], float(np.abs(X[b] - Xr).max()))
        worst["U"] = max(worst["U"], float(np.abs(U[b] - Ur).max()))
    print(f"{config}: verdict and iteration counts of the reference build on {same}/{B} instances; worst differences "
          f"among those {worst}")
    assert same >= 0.9 * B
    assert worst["cost"] < 1e-8 and worst["X"] < 1e-6 and worst["U"] < 5e-6 and worst["viol"] < 1e-8


def ref_generic(lib, spec, x0):
    """any ProblemSpec on the reference's solver (oracle/ref_shim/ref_driver.cpp replays its builder calls)"""
    if not hasattr(lib, "altro_refb_solve"):
        pytest.skip("this build of oracle/_ref has no generic builder")
    handle = spec.build(lib, "altro_refb_")
    n, m, N = spec.n, spec.m, spec.N
    X = np.zeros((N + 1, n)); U = np.zeros((N, m)); sc = np.zeros(4); it = np.zeros(4, dtype=np.int32)
    U0 = np.ascontiguousarray(spec.initial_controls(), dtype=np.float64)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.altro_refb_solve.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 7
    got = lib.altro_refb_solve(handle, 1, ptr(x0), ptr(U0), None, ptr(X), ptr(U), ptr(sc), ptr(it))
    lib.altro_refb_problem_destroy.argtypes = [ctypes.c_void_p]
    lib.altro_refb_problem_destroy(handle)
    assert got == N
    return X, U, sc, it


_generic_solved = {}


@pytest.mark.parametrize("config", ["cartpole-c4", "random-lqr-c5"])
def test_device_solves_follow_the_reference_solver_on_this_repos_models(gpu, ref, config):
    """C4 and C5: models the reference does not have, given to its solver as user functors (the closed forms of the
    oracle).  Tolerances are those of tests/test_gpu_parity.py for these configurations: 100 iterations of a swing-up
    and an n = 32 Riccati recursion amplify the last bits."""
    if config == "cartpole-c4":
        spec = P.cartpole_problem(N=200)
        X0 = P.perturbed_initial_states(spec, 16, P.CARTPOLE_X0_SCALE)
        tol = dict(cost=1e-8, X=1e-7, U=1e-6)
    else:
        spec = P.random_lqr_problem()
        X0 = P.normal_initial_states(spec, 16)
        tol = dict(cost=1e-9, X=1e-6, U=1e-5)
    B = X0.shape[0]
    s = gpu.BatchSolver(spec, B)
    s.set_inputs(X0)
    s.solve_al()
    res = s.results()
    X, U = s.trajectory()
    same, worst = 0, dict(cost=0.0, X=0.0, U=0.0)
    for b in range(B):
        key = (config, b)
        if key not in _generic_solved:
            _generic_solved[key] = ref_generic(ref, spec, X0[b])
        Xr, Ur, sc, it = _generic_solved[key]
        if (int(res["status"][b]), int(res["iters"][b, 1]), int(res["iters"][b, 2])) != (int(it[0]), int(it[2]), int(it[3])):
            continue
        same += 1
        worst["cost"] = max(worst["cost"], abs(res["cost"][b] - sc[0]) / max(1.0, abs(sc[0])))
        worst["X"] = max(worst["X"], float(np.abs(X[b] - Xr).max()))
        worst["U"] = max(worst["U"], float(np.abs(U[b] - Ur).max()))
    print(f"{config}: verdict and iteration counts of the reference's solver on {same}/{B} instances; worst differences "
          f"among those {worst}")
    assert same >= 0.9 * B
    assert all(worst[k] <= tol[k] for k in tol), (worst, tol)
