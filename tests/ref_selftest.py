"""The control experiment behind tests/test_reference_unit_tests.py: the reference's OWN unit tests against the
reference's OWN code (headers and sources read where they lie under /root/reference), built with the same two
stand-ins the mirror runs use — this repo's Eigen stand-in and the GoogleTest stand-in.  If the reference passes its
own tests on them, a failure of the mirror + device runs is a property of the mirror or the device, not of the test
scaffolding; and where the reference itself fails an assertion on them (the two 4-ulp cost assertions of
auglag_test.cpp), that assertion tests Eigen's summation order rather than the algorithm.

Executables go to a temporary directory; nothing travels, nothing of the reference is copied into the repository.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(tempfile.gettempdir(), "altro_b200_ref_selftest")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


build_ref = _load("build_ref", os.path.join(ROOT, "oracle", "build_ref.py"))
ref_unit = _load("ref_unit_tests", os.path.join(ROOT, "tests", "ref_unit_tests.py"))
TESTS = ref_unit.HOST_TESTS + ref_unit.DEVICE_TESTS


def available():
    return build_ref.available() and os.path.isdir(os.path.join(REF, "test"))


def _flags():
    # /root/reference first: every altro/... header is the reference's; only eigen3/Eigen/Dense and gtest/gtest.h are ours
    return ["g++", "-std=c++14", "-O1", "-w", "-DFMT_HEADER_ONLY",
            "-include", os.path.join(ROOT, "oracle", "ref_shim", "compat.hpp"),
            "-I", REF, "-I", os.path.join(ROOT, "altro_cpp_b200", "host", "include"), "-I", ref_unit.STANDIN,
            "-I", build_ref.fmt_include(), f'-DLOCAL_LOG_DIR="{OUT}"', f'-DLOGDIR="{OUT}"']   # no -DNDEBUG: death tests


def _compile_one(src):
    obj = os.path.join(OUT, os.path.relpath(src, "/").replace("/", "_") + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj
    r = subprocess.run(_flags() + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise subprocess.CalledProcessError(r.returncode, src, output=r.stdout + r.stderr)
    return obj


def reference_objects(jobs=8):
    """the reference's solver and example sources compiled (assertions on) for linking into test programs"""
    os.makedirs(OUT, exist_ok=True)
    sources = [s for s in build_ref.sources() if not s.endswith("ref_driver.cpp")]
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        return list(pool.map(_compile_one, sources))


def build_program(src, name):
    """one program of this repo (a probe) against the reference's own headers and sources -> executable path"""
    exe = os.path.join(OUT, name)
    subprocess.check_call(_flags() + [src] + reference_objects() + ["-o", exe, "-lpthread"])
    return exe


def build_all(jobs=8):
    """-> {rel: (exe or None, log)}"""
    os.makedirs(OUT, exist_ok=True)
    sources = [s for s in build_ref.sources() if not s.endswith("ref_driver.cpp")]
    sources.append(os.path.join(ref_unit.STANDIN, "gtest_main.cc"))
    compile_one = _compile_one

    def link_one(rel):
        exe = os.path.join(OUT, rel.replace("/", "_").replace(".cpp", ""))
        r = subprocess.run(_flags() + [os.path.join(REF, "test", rel)] + objs + ["-o", exe, "-lpthread"],
                           capture_output=True, text=True)
        return (exe if r.returncode == 0 else None), r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=jobs) as pool:
        objs = list(pool.map(compile_one, sources))
        results = list(pool.map(link_one, TESTS))
    return dict(zip(TESTS, results))


if __name__ == "__main__":
    if not available():
        sys.exit("the reference sources are not mounted here")
    for rel, (exe, log) in build_all().items():
        if exe is None:
            print("COMPILE-FAIL", rel, [l for l in log.splitlines() if "error" in l][:3])
            continue
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        summary = [l for l in r.stdout.splitlines() if l.startswith("[====")]
        print("PASS" if r.returncode == 0 else "FAIL", rel, summary[-1] if summary else "")
        for l in r.stdout.splitlines():
            if "FAILED" in l:
                print("    ", l)
