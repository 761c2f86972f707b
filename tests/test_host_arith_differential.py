"""The host mirror's single-knot arithmetic against the reference's own: tests/cpp/host_arith_probe.cpp (ALCost value /
gradient / Hessian with multipliers and penalties at an interior and at the terminal knot, dual and penalty updates,
KnotPointFunctions::CalcActionValueExpansion / CalcGains / CalcCostToGo) is built twice from the same source — against
altro_cpp_b200/host/include + libaltro_b200.so, and against /root/reference's headers and sources (on the Eigen
stand-in) — and the printed numbers are compared.  They agree to the last one or two bits (the mirror sums some
products in another order); the bound asserted is 1e-13 relative.  CPU only; needs /root/reference."""
import importlib.util
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


NUMBER = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")


def test_mirror_host_arithmetic_matches_the_reference():
    selftest = _load("ref_selftest")
    if not selftest.available():
        pytest.skip("the reference sources are not mounted here")
    ref_unit = selftest.ref_unit
    import altro_cpp_b200 as pkg
    pkg.lib()
    src = os.path.join(ROOT, "tests", "cpp", "host_arith_probe.cpp")
    support = [o for o in ref_unit.build_support() if not o.endswith("gtest_main.o")]
    mirror_exe = os.path.join(ref_unit.OUT, "host_arith_probe")
    subprocess.check_call(ref_unit._flags() + [src] + support + ["-o", mirror_exe, "-L", ref_unit.LIBDIR, "-laltro_b200",
                                                                 f"-Wl,-rpath,{ref_unit.LIBDIR}", "-lpthread"])
    reference_exe = selftest.build_program(src, "host_arith_probe")
    mine = subprocess.run([mirror_exe], capture_output=True, text=True, check=True).stdout.splitlines()
    theirs = subprocess.run([reference_exe], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(mine) == len(theirs) and len(mine) > 90
    worst, compared = 0.0, 0
    for a, b in zip(mine, theirs):
        assert NUMBER.sub("#", a) == NUMBER.sub("#", b), (a, b)   # same labels and shapes
        if a.startswith(" "):  # a row of coefficients
            for x, y in zip(map(float, NUMBER.findall(a)), map(float, NUMBER.findall(b))):
                err = abs(x - y) / max(1.0, abs(y))
                worst = max(worst, err)
                compared += 1
                assert err <= 1e-13, (a, b)
        else:
            assert a.split()[:-1] == b.split()[:-1]
            xs, ys = NUMBER.findall(a.split(" ", 1)[1] if " " in a else a), NUMBER.findall(b.split(" ", 1)[1] if " " in b else b)
            for x, y in zip(map(float, xs), map(float, ys)):
                assert abs(x - y) <= 1e-13 * max(1.0, abs(y)), (a, b)
    print(f"{compared} coefficients compared, worst relative difference {worst:.2e}")
    assert compared >= 200
