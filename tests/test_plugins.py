"""Plug-in dynamics models: a model is the CUDA source of a functor struct; the
library compiles its kernels for it with NVRTC, caches the module and runs it like a built-in one.
Reference analogue: subclassing ContinuousDynamics (altro/problem/dynamics.hpp:59-99), e.g.
examples/unicycle.hpp."""
import os

import numpy as np
import pytest

import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the unicycle of examples/unicycle.cpp:12-33 written against the model concept of csrc/device.cuh, without
# the shared-stage shortcut of the built-in functor (kStage3RepeatsStage2 = false evaluates stage 3 itself:
# same values, since stages 2 and 3 of this model coincide bit for bit)
UNICYCLE_PLUGIN = r"""
struct PluginUnicycle {
  static constexpr int n = 3, m = 2;
  static constexpr bool kDiscrete = false;
  static constexpr bool kStage3RepeatsStage2 = false;
  static __device__ __forceinline__ void eval(const double*, const double* x, const double* u, double* xd) {
    double s, c;
    sincos(x[2], &s, &c);
    xd[0] = u[0] * c;
    xd[1] = u[0] * s;
    xd[2] = u[1];
  }
  static __device__ __forceinline__ void jac(const double*, const double* x, const double* u, double* A, double* B) {
    double s, c;
    sincos(x[2], &s, &c);
    for (int i = 0; i < 9; ++i) A[i] = 0.0;
    for (int i = 0; i < 6; ++i) B[i] = 0.0;
    A[0 + 2 * 3] = -u[0] * s;
    A[1 + 2 * 3] = u[0] * c;
    B[0 + 0 * 3] = c;
    B[1 + 0 * 3] = s;
    B[2 + 1 * 3] = 1.0;
  }
};
"""


def test_register_and_precompile_without_a_gpu(tmp_path, monkeypatch):
    monkeypatch.setenv("ALTRO_B200_MODULE_CACHE", str(tmp_path))
    mid = pkg.register_model("PluginUnicycle", UNICYCLE_PLUGIN, 3, 2, 0)
    assert mid >= 100
    assert pkg.register_model("PluginUnicycle", UNICYCLE_PLUGIN, 3, 2, 0) == mid  # idempotent
    assert pkg.lib().altro_b200_is_supported(3, 2, mid) == 1 and pkg.lib().altro_b200_is_supported(4, 2, mid) == 0
    path = pkg.precompile_model(mid)
    assert os.path.exists(path) and os.path.getsize(path) > 100_000 and path.startswith(str(tmp_path))
    with pytest.raises(pkg.SolverError, match="already registered"):
        pkg.register_model("PluginUnicycle", UNICYCLE_PLUGIN + " ", 3, 2, 0)


def test_a_model_that_does_not_compile_is_reported(tmp_path, monkeypatch):
    monkeypatch.setenv("ALTRO_B200_MODULE_CACHE", str(tmp_path))
    bad = "struct Broken { static constexpr int n = 2, m = 1; static __device__ void eval(const double*, const double*, const double*, double* xd) { xd[0] = undefined_symbol; } };"
    mid = pkg.register_model("Broken", bad, 2, 1, 0)
    with pytest.raises(pkg.SolverError, match="does not compile"):
        pkg.precompile_model(mid)


def test_shipped_cartpole_plugin_is_in_the_module_cache():
    """__graft_entry__.build() compiles plugins/cartpole.cuh ahead of time: the GPU tests load it from the cache."""
    path = pkg.precompile_model(P.MODEL_CARTPOLE)
    assert os.path.basename(path).startswith("Cartpole_") and os.path.exists(path)


def _spec_with_model(model_id):
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    spec.calls = [(c if c[0] != "set_model" else ("set_model", model_id) + tuple(c[2:])) for c in spec.calls]
    return spec


@pytest.mark.gpu
def test_plugin_unicycle_matches_the_builtin_bit_for_bit():
    """The same model once compiled into the library and once supplied as source at run time: every
    output of a C2-style batch is identical (the kernels are the same templates)."""
    import torch
    assert torch.cuda.is_available()
    mid = pkg.register_model("PluginUnicycle", UNICYCLE_PLUGIN, 3, 2, 0)
    B = 300
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    X0 = P.perturbed_initial_states(spec, B, P.UNICYCLE_X0_SCALE)
    a = pkg.BatchSolver(spec, B).solve_al_host(X0)
    b = pkg.BatchSolver(_spec_with_model(mid), B).solve_al_host(X0)
    for k in ("status", "iters", "cost", "viol", "X", "U"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_cartpole_runs_as_a_plugin(oracle):
    """BASELINE config C4's model is not compiled into the library any more: it is plugins/cartpole.cuh."""
    spec = P.cartpole_problem(N=200)
    X0 = P.perturbed_initial_states(spec, 32, P.CARTPOLE_X0_SCALE)
    out = pkg.BatchSolver(spec, 32).solve_al_host(X0)
    ref = oracle.solve_batch(spec, X0, nthreads=8, want_gains=False)
    assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["status"], ref["status"])
    assert np.abs(out["X"] - ref["X"]).max() <= 1e-6
    so = os.path.join(ROOT, "altro_cpp_b200", "libaltro_b200.so")
    assert open(so, "rb").read().find(b"NS_8CartpoleE") < 0  # no kernel instantiated on it inside the library


@pytest.mark.gpu
def test_triple_integrator_with_one_degree_of_freedom_is_instantiated_at_run_time():
    """ALTRO_B200_MODEL_TRIPLE_INTEGRATOR with dof = 1 (n = 3, m = 1) is not compiled into the library: the
    TripleIntegrator<dof> template of device.cuh is instantiated by NVRTC (module pre-compiled by build()).
    Constrained AL solve against the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (the product path has no CPU fallback)")
    import altro_cpp_b200 as pkg
    from altro_cpp_b200 import problems as P
    from oracle import binding as ob
    pkg.set_default_engine(None)
    spec = P.triple_integrator_problem(dof=1, N=30, add_constraints=True)
    B = 40
    X0 = P.perturbed_initial_states(spec, B, np.asarray(P.TRIPLE_INTEGRATOR_X0_SCALE)[[0, 2, 4]])
    out = pkg.BatchSolver(spec, B).solve_al_host(X0)
    ref = ob.solve_batch(spec, X0, nthreads=4, want_gains=False)
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["iters"], ref["iters"])
    assert np.abs(out["X"] - ref["X"]).max() <= 1e-8 * max(1.0, np.abs(ref["X"]).max())
    assert np.abs(out["cost"] - ref["cost"]).max() <= 1e-9 * max(1.0, np.abs(ref["cost"]).max())
