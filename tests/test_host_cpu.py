"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol declared in
include/altro_b200.h, the problem builder validates arguments like the reference's asserts, and
the product path fails loudly — it never falls back to a CPU implementation."""
import ctypes
import os
import re

import numpy as np
import pytest

import altro_cpp_b200 as pkg
from altro_cpp_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "altro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(altro_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = pkg.lib()
    names = declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_support_matrix():
    L = pkg.lib()
    assert b"sm_100a" in L.altro_b200_version()
    assert L.altro_b200_is_supported(3, 2, P.MODEL_UNICYCLE) == 1
    assert L.altro_b200_is_supported(6, 2, P.MODEL_TRIPLE_INTEGRATOR) == 1
    assert L.altro_b200_is_supported(4, 1, P.MODEL_CARTPOLE) == 1
    assert L.altro_b200_is_supported(32, 8, P.MODEL_LINEAR) == 1
    assert L.altro_b200_is_supported(5, 5, P.MODEL_UNICYCLE) == 0


def test_default_options_match_reference():
    # altro/common/solver_options.hpp:23-56
    o = pkg.default_options()
    assert (o.max_iterations_total, o.max_iterations_outer, o.max_iterations_inner) == (300, 30, 100)
    assert (o.cost_tolerance, o.gradient_tolerance) == (1e-4, 1e-2)
    assert (o.bp_reg_increase_factor, o.bp_reg_initial, o.bp_reg_max, o.bp_reg_min) == (1.6, 0.0, 1e8, 1e-8)
    assert (o.bp_reg_fail_threshold, o.check_forwardpass_bounds, o.line_search_max_iterations) == (100, 1, 20)
    assert (o.line_search_lower_bound, o.line_search_upper_bound, o.line_search_decrease_factor) == (1e-8, 10.0, 2.0)
    assert (o.constraint_tolerance, o.maximum_penalty, o.initial_penalty, o.reset_duals) == (1e-4, 1e8, 1.0, 1)
    assert o.penalty_scaling == 10.0


def test_problem_builder_validates_arguments():
    L = pkg.lib()
    h = ctypes.c_void_p()
    assert L.altro_b200_problem_create(3, 2, 0, ctypes.byref(h)) == -1          # N must be positive
    assert L.altro_b200_problem_create(3, 2, 10, ctypes.byref(h)) == 0
    dp = ctypes.POINTER(ctypes.c_double)
    lb = np.array([1.0, 0.0]); ub = np.array([0.0, 1.0])
    L.altro_b200_problem_add_control_bound.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp]
    rc = L.altro_b200_problem_add_control_bound(h, 0, lb.ctypes.data_as(dp), ub.ctypes.data_as(dp))
    assert rc == -1 and b"Lower bound" in L.altro_b200_last_error()               # basic_constraints.hpp:132-135
    assert L.altro_b200_problem_add_control_bound(h, 11, ub.ctypes.data_as(dp), lb.ctypes.data_as(dp)) == -1
    L.altro_b200_problem_destroy(h)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SolverError, match="no usable CUDA device"):
        pkg.BatchSolver(P.unicycle_problem(), 8)


def test_product_does_not_reference_the_oracle():
    """Nothing under altro_cpp_b200/ or include/ may import, include or link oracle/."""
    bad = []
    for base in ("altro_cpp_b200", "include"):
        for dp_, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp_, f)).read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|altro_oracle_[a-z]+\(", txt):
                        if f == "problems.py":  # docstring mentions the prefix only
                            txt2 = re.sub(r'""".*?"""', "", txt, flags=re.S)
                            if not re.search(r"(from|import)\s+oracle|liboracle", txt2):
                                continue
                        bad.append(os.path.join(dp_, f))
    assert not bad, bad


def test_batch_generator_is_deterministic_and_counter_based():
    spec = P.unicycle_problem(P.K_THREE_OBSTACLES)
    a = P.perturbed_initial_states(spec, 64, P.UNICYCLE_X0_SCALE)
    b = P.perturbed_initial_states(spec, 32, P.UNICYCLE_X0_SCALE, first=32)
    assert np.array_equal(a[32:], b)          # stream = instance index: shards agree with the whole
    assert np.array_equal(a[0], spec.x0)      # instance 0 is the nominal problem
    assert np.all(np.abs(a[:, :2]) <= 0.3) and np.all(np.abs(a[:, 2]) <= np.pi / 6)
