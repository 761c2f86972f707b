"""The reference's OWN unit tests (test/**/*.cpp of optimusride/altro-cpp), unmodified, against this repo's host
mirror: compiled where they lie under /root/reference with the mirror's headers, the Eigen stand-in and a small
GoogleTest stand-in (tests/cpp/gtest_standin), linked with libaltro_b200.so (tests/ref_unit_tests.py).

* host-side classes (KnotPoint, Problem, cost / dynamics functors, derivative checks, expansions, thread pool, timer,
  logger, options): 16 test programs, run here on the CPU;
* solver tests (unicycle_ilqr_test, ilqr_test, ilqr_class_test, example_unicycle_test, example_triple_integrator_test): built
  here, run on the GPU box from the executables that travel with the snapshot — the reference's golden iteration
  counts, costs, step lengths and gains, asserted by the reference's own code, on the device.
"""
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
    tail = "\n".join(l for l in (r.stdout + r.stderr).splitlines() if "Failure" in l or "FAILED" in l or "Expected" in l or "actual" in l)[-3000:]
    assert r.returncode == 0, tail
    assert " 0 failed." in r.stdout
