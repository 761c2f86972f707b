"""Control experiment for tests/test_reference_unit_tests.py (see tests/ref_selftest.py): the reference's own unit
tests against the reference's own code, on this repo's Eigen and GoogleTest stand-ins.  161 of its 163 cases pass; the
two that do not are the 4-ulp cost assertions of auglag_test.cpp — the same two, with the same value to 1e-15, that the
device run reports.  CPU only; needs /root/reference."""
import importlib.util
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("ref_selftest", os.path.join(ROOT, "tests", "ref_selftest.py"))
selftest = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(selftest)


@pytest.fixture(scope="module")
def built():
    if not selftest.available():
        pytest.skip("the reference sources are not mounted here")
    return selftest.build_all()


@pytest.mark.timeout(1200)
def test_the_reference_builds_against_its_own_tests_on_the_standins(built):
    failed = {rel: log[-1200:] for rel, (exe, log) in built.items() if exe is None}
    assert not failed, failed


@pytest.mark.parametrize("rel", selftest.TESTS)
def test_the_reference_passes_its_own_unit_test_on_the_standins(built, rel):
    exe, log = built[rel]
    assert exe is not None, log[-1200:]
    for attempt in range(3 if "timer_test" in rel else 1):  # timer_test bounds a wall-clock overhead
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        if r.returncode == 0:
            break
    ulp = selftest.ref_unit.ULP_ASSERTIONS.get(rel)
    if ulp is None:
        assert r.returncode == 0, (r.stdout + r.stderr)[-2500:]
        assert " 0 failed." in r.stdout
        return
    # auglag_test.cpp: EXPECT_DOUBLE_EQ(cost, 0.03893465058924039) was generated with the real Eigen; with any other
    # summation order — the stand-in's here, the device's in the GPU run — the reference's algorithm lands 1.5e-12 away
    text = r.stdout + r.stderr
    where = re.findall(r"^(\S+):(\d+): Failure\n(.*)$", text, flags=re.M)
    assert r.returncode != 0 and len(where) == 2 and "unexpected exception" not in text, text[-2500:]
    for path, line, detail in where:
        assert path.endswith(rel) and int(line) in ulp["lines"]
        got = float(re.search(r": (\S+) vs ", detail).group(1))
        assert abs(got - ulp["golden"]) / ulp["golden"] < 1e-11
        assert abs(got - 0.03893465058918357) / ulp["golden"] < 1e-14   # what the B200 returns (profiles/r02_reference_auglag_test_gpu.log)
    assert "15 tests ran, 2 failed." in r.stdout
