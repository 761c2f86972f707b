"""The drop-in claim, checked literally: the reference's own examples/*.cpp, examples/problems/*.cpp
and perf/*.cpp — unmodified, read from /root/reference — compile against this repo's altro/ headers
and link with the device library (north_star: "examples/ and perf/ compile unchanged").

Their classes (Unicycle, QuadraticCost, ControlBound, CircleConstraint ...) know nothing about the
device: the solver recognises them through their virtual interface (altro/device_registry.hpp).
On a GPU the reference's perf/benchmark_unicycle.cpp then solves its problem on the device with
the iteration count the reference's own tests expect."""
import os
import subprocess

import pytest

import altro_cpp_b200 as pkg

import importlib.util
_spec = importlib.util.spec_from_file_location("ref_build", os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_build.py"))
ref_build = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_build)


@pytest.fixture(scope="module")
def programs():
    if not ref_build.available():
        pytest.skip("/root/reference (or the fmt headers its sources include) is not present on this machine")
    pkg.lib()  # makes sure libaltro_b200.so exists
    try:
        return ref_build.build()
    except subprocess.CalledProcessError as e:
        pytest.fail(f"{e.cmd} does not build against the host mirror:\n{e.output[-3000:]}")


def has_gpu():
    import torch
    return torch.cuda.is_available()


def test_reference_examples_and_perf_build_unchanged(programs):
    assert sorted(programs) == sorted(ref_build.PROGRAMS)
    for exe in programs.values():
        assert os.path.exists(exe)


def test_reference_threadpool_benchmark_runs(programs):
    r = subprocess.run([programs["benchmark_threadpool"]], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(os.path.join(ref_build.OUT, "profile_threadpool.out")).read()
    assert "Number of threads" in out and "Speedup" in out


def test_reference_benchmark_is_loud_without_gpu(programs):
    if has_gpu():
        pytest.skip("a GPU is present")
    r = subprocess.run([programs["benchmark_unicycle"], "2"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr  # altro::DeviceError escapes main(): there is no CPU fallback


def _prebuilt(name):
    exe = os.path.join(ref_build.OUT, name)
    if not os.path.exists(exe):
        pytest.skip(f"tests/_ref_build/{name} was not built (it is built where /root/reference exists and travels "
                    "with the snapshot)")
    return exe


@pytest.mark.gpu
def test_reference_benchmark_unicycle_runs_on_device():
    """perf/benchmark_unicycle.cpp of the reference, unmodified: 3-obstacle problem, two solves in a loop.
    Expected by the reference's tests (and reproduced by the CPU oracle): 50 iLQR iterations."""
    r = subprocess.run([_prebuilt("benchmark_unicycle"), "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("iters = 50") == 2, r.stdout


@pytest.mark.gpu
def test_reference_benchmark_triple_integrator_runs_on_device():
    r = subprocess.run([_prebuilt("benchmark_triple_integrator")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Total Compute Time") == 2


@pytest.mark.gpu
def test_reference_benchmark_expansions_runs_on_device():
    r = subprocess.run([_prebuilt("benchmark_expansions")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = open(os.path.join(ref_build.OUT, "profile_expansions.out")).read()
    assert "Serial time" in out and "tasks" in out
