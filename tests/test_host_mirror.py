"""The host mirror of the reference's C++ API (altro_cpp_b200/host/include): reference-style programs
compile against it, refuse to run without a GPU and reproduce the reference's golden numbers on
one (tests/cpp/host_mirror_test.cpp)."""
import os
import shutil
import subprocess

import pytest

import altro_cpp_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "altro_cpp_b200")
HOST = os.path.join(LIBDIR, "host", "include")          # altro/ + the Eigen stand-in: the library proper
EXAMPLES = os.path.join(LIBDIR, "host", "examples_b200")  # this repo's own examples/ and perf/ programs


def compile_program(tmp_path, src, name):
    pkg.lib()  # makes sure libaltro_b200.so exists
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror",
                           "-I", os.path.join(ROOT, "include"), "-I", HOST, "-I", EXAMPLES, src, "-o", exe,
                           "-L", LIBDIR, "-laltro_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def has_gpu():
    import torch
    return torch.cuda.is_available()


def test_host_mirror_cpu_checks_and_loud_failure(tmp_path):
    exe = compile_program(tmp_path, os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "host_mirror_test")
    if has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe, "cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpu: 0 failure(s)" in r.stdout


def test_benchmark_programs_refuse_to_run_without_gpu(tmp_path):
    if has_gpu():
        pytest.skip("a GPU is present")
    for name in ("benchmark_unicycle", "benchmark_triple_integrator"):
        exe = compile_program(tmp_path, os.path.join(EXAMPLES, "perf", name + ".cpp"), name)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2
        assert "no usable CUDA device" in r.stderr


@pytest.mark.skipif(shutil.which("cmake") is None, reason="cmake not installed")
def test_cmake_project_with_reference_target_names(tmp_path):
    pkg.lib()
    build = str(tmp_path / "build")
    subprocess.check_call(["cmake", "-S", ROOT, "-B", build, "-DALTRO_B200_PREBUILT=ON"],
                          stdout=subprocess.DEVNULL)
    subprocess.check_call(["cmake", "--build", build, "-j", "4"], stdout=subprocess.DEVNULL)
    for exe in ("benchmark_unicycle", "benchmark_triple_integrator", "host_mirror_test"):
        assert os.path.exists(os.path.join(build, exe))
    # a downstream project that links the reference's target name
    down = tmp_path / "downstream"
    down.mkdir()
    (down / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.18)\nproject(Down LANGUAGES CXX)\n"
        f"set(ALTRO_B200_PREBUILT ON CACHE BOOL \"\")\nset(ALTRO_BUILD_TESTS OFF CACHE BOOL \"\")\n"
        f"set(ALTRO_BUILD_BENCHMARKS OFF CACHE BOOL \"\")\n"
        f"add_subdirectory({ROOT} altro)\nadd_executable(app app.cpp)\n"
        "target_link_libraries(app PRIVATE altro::altro altro::example_problems)\n")
    (down / "app.cpp").write_text(
        '#include "altro/augmented_lagrangian/al_solver.hpp"\n#include "examples/problems/unicycle.hpp"\n'
        "int main() { altro::problems::UnicycleProblem p; return p.MakeProblem().NumSegments() == 100 ? 0 : 1; }\n")
    subprocess.check_call(["cmake", "-S", str(down), "-B", str(down / "b")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["cmake", "--build", str(down / "b")], stdout=subprocess.DEVNULL)
    assert subprocess.run([str(down / "b" / "app")]).returncode == 0


@pytest.mark.gpu
def test_host_mirror_reproduces_reference_goldens_on_device(tmp_path):
    exe = compile_program(tmp_path, os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "host_mirror_test")
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu: 0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_benchmark_unicycle_single_and_batched(tmp_path):
    exe = compile_program(tmp_path, os.path.join(EXAMPLES, "perf", "benchmark_unicycle.cpp"), "benchmark_unicycle")
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("iters = 50, outer = 5, status = 0") == 2
    r = subprocess.run([exe, "2", "512"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("nominal iters = 50") == 2


def _mirror_headers():
    out = []
    for d, _, files in os.walk(os.path.join(HOST, "altro")):
        out += [os.path.relpath(os.path.join(d, f), HOST) for f in files if f.endswith(".hpp")]
    return sorted(out)


def test_every_mirror_header_is_self_contained():
    """`#include "altro/<any header>"` alone compiles (no header relies on what another one happened to include)."""
    from concurrent.futures import ThreadPoolExecutor

    def check(rel):
        r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", HOST,
                            "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                           input=f'#include "{rel}"\nint main() {{ return 0; }}\n', capture_output=True, text=True)
        return rel, r.returncode, r.stderr[-800:]

    with ThreadPoolExecutor(max_workers=8) as pool:
        failed = [(rel, err) for rel, rc, err in pool.map(check, _mirror_headers()) if rc != 0]
    assert not failed, failed


def test_mirror_covers_the_reference_header_tree():
    ref = "/root/reference/altro"
    if not os.path.isdir(ref):
        pytest.skip("the reference sources are not mounted here")
    mine = set(_mirror_headers())
    missing = []
    for d, _, files in os.walk(ref):
        for f in files:
            rel = os.path.relpath(os.path.join(d, f), os.path.dirname(ref))
            if f.endswith(".hpp") and rel not in mine:
                missing.append(rel)
    assert not missing, missing


def test_mirror_offers_the_public_methods_of_the_reference_headers():
    """Every method the reference's headers declare in a public section has a counterpart of the same name in the
    mirror header of the same path.  A name-level audit (signatures are exercised by compiling the reference's own
    programs and tests): it catches a method that was never offered."""
    import re
    ref = "/root/reference/altro"
    if not os.path.isdir(ref):
        pytest.skip("the reference sources are not mounted here")
    decl = re.compile(r"^  (?:template\s*<[^>]*>\s*)?(?:static\s+|virtual\s+|explicit\s+|inline\s+|constexpr\s+)*"
                      r"[A-Za-z_:<>,&\*\s]*?\b([A-Z][A-Za-z0-9]*)\s*\(")
    # calls inside inline bodies that the line-based scan picks up, private helpers, and one method the reference
    # declares but never defines (FunctionBase::TestCheck)
    not_api = {"TestCheck", "DefaultLogger", "SetData", "Init", "CalcIndividualCosts", "DecreaseRegularization",
               "IncreaseRegularization"}
    # public in the reference, deliberately not offered: one line-search candidate on the host-visible
    # buffer Zbar_, and the device half of SolveSetup()
    not_offered = {"ilqr/ilqr.hpp": {"RolloutClosedLoop", "ResetInternalVariables"}}
    missing = {}
    for d, _, files in os.walk(ref):
        for f in files:
            if not f.endswith(".hpp"):
                continue
            rel = os.path.relpath(os.path.join(d, f), ref)
            mirror = open(os.path.join(HOST, "altro", rel)).read()
            public = False
            names = set()
            for line in open(os.path.join(d, f)):
                if re.match(r"^ public:", line):
                    public = True
                elif re.match(r"^ (private|protected):", line):
                    public = False
                elif re.match(r"^(class|struct)\b", line):
                    public = line.startswith("struct")
                elif public and not line.strip().startswith(("//", "*", "return", "ALTRO_")):
                    m = decl.match(line)
                    if m:
                        names.add(m.group(1))
            gone = sorted(n for n in names - not_api - not_offered.get(rel, set()) if not re.search(r"\b" + n + r"\b", mirror))
            if gone:
                missing[rel] = gone
    assert not missing, missing
