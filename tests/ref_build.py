"""Compiles the UNMODIFIED sources of the reference's examples/ and perf/ directories (read where
they lie, /root/reference) against this repo's host mirror (altro_cpp_b200/host/include: altro/ headers +
the Eigen stand-in) and links them with libaltro_b200.so.  Nothing of the reference is copied: only
the resulting executables are kept, under tests/_ref_build/ (git-ignored; they travel to the GPU
box, where /root/reference does not exist, so that the -m gpu tests can run them on the device).
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "_ref_build")
LIBDIR = os.path.join(ROOT, "altro_cpp_b200")

EXAMPLE_SOURCES = ["examples/unicycle.cpp", "examples/triple_integrator.cpp", "examples/quadratic_cost.cpp",
                   "examples/basic_constraints.cpp", "examples/obstacle_constraints.cpp",
                   "examples/problems/unicycle.cpp", "examples/problems/triple_integrator.cpp"]
PROGRAMS = ["benchmark_unicycle", "benchmark_triple_integrator", "benchmark_expansions", "benchmark_threadpool"]


def fmt_include():
    """fmt is a dependency of the reference's sources (not of this repo): torch ships its headers."""
    try:
        import torch
        inc = os.path.join(os.path.dirname(torch.__file__), "include")
        return inc if os.path.exists(os.path.join(inc, "fmt", "format.h")) else None
    except Exception:
        return None


def available():
    return os.path.isdir(os.path.join(REF, "perf")) and fmt_include() is not None


def ref_include_dir():
    """An include root that shows the reference's examples/, perf/ and test/ trees but NOT its altro/ headers, so
    that `#include "altro/..."` can only ever resolve to this repo's mirror (a header the mirror lacks is a
    compile error, not a silent fall-back to the reference's implementation)."""
    import tempfile
    inc = os.path.join(tempfile.gettempdir(), "altro_b200_ref_include")  # links only; kept out of the snapshot
    os.makedirs(inc, exist_ok=True)
    for name in os.listdir(REF):
        src, dst = os.path.join(REF, name), os.path.join(inc, name)
        if name == "altro" or not os.path.isdir(src) or name.startswith("."):
            continue
        if not os.path.islink(dst):
            os.symlink(src, dst)
    return inc


def build(verbose=False):
    """-> {program: path}.  Raises CalledProcessError (with the compiler output) on any failure."""
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "altro_cpp_b200", "host", "include"), "-I", os.path.join(ROOT, "include"),
           "-I", ref_include_dir(), "-I", fmt_include(), "-DFMT_HEADER_ONLY", f'-DLOCAL_LOG_DIR="{OUT}"']
    flags = ["g++", "-std=c++14", "-O1", "-DNDEBUG"]  # the reference's CI builds Release
    objs = []
    procs = []
    for src in EXAMPLE_SOURCES + [f"perf/{p}.cpp" for p in PROGRAMS]:
        obj = os.path.join(OUT, src.replace("/", "_") + ".o")
        procs.append((src, obj, subprocess.Popen(flags + inc + ["-c", os.path.join(REF, src), "-o", obj],
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise subprocess.CalledProcessError(p.returncode, src, output=log)
        if verbose and log.strip():
            print(log)
        if src.startswith("examples/"):
            objs.append(obj)
    out = {}
    for prog in PROGRAMS:
        exe = os.path.join(OUT, prog)
        cmd = ["g++", "-o", exe, os.path.join(OUT, f"perf_{prog}.cpp.o")] + objs + \
              ["-L", LIBDIR, "-laltro_b200", f"-Wl,-rpath,{LIBDIR}", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, prog, output=r.stdout + r.stderr)
        out[prog] = exe
    for f in os.listdir(OUT):
        if f.endswith(".o"):
            os.remove(os.path.join(OUT, f))
    return out
