"""Builds oracle/_ref/libaltro_ref.so: the REFERENCE's own solver sources (altro/**/*.cpp, examples/*.cpp,
examples/problems/*.cpp — compiled where they lie under /root/reference, nothing copied) plus the small C entry
point oracle/ref_shim/ref_driver.cpp.

The reference depends on Eigen, which this image does not have; its own build (cmake + find_package(Eigen3)) cannot
run here and DESIGN.md records the reference as not buildable as shipped.  This recipe compiles it against the Eigen
stand-in of the host mirror instead (altro_cpp_b200/host/include/eigen3/Eigen/Dense) — include order puts
/root/reference first, so every `altro/...` header is the reference's and only `eigen3/Eigen/Dense` comes from this
repo — and force-includes oracle/ref_shim/compat.hpp (fmt formatter specialisations for the newer fmt of this image).
Left out: altro/main.cpp (a demo main).

The result is the reference's control flow and formulas on the stand-in's linear algebra: a cross-check of the oracle
restatement's logic (status, iteration counts, line-search decisions, AL updates), not of Eigen's rounding order.
Used only by tests/ (never by the product, never by bench.py's timed paths).

    python oracle/build_ref.py        # -> oracle/_ref/libaltro_ref.so
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(OUT, "libaltro_ref.so")
SKIP = {"altro/main.cpp"}


def fmt_include():
    try:
        import torch
        inc = os.path.join(os.path.dirname(torch.__file__), "include")
        return inc if os.path.exists(os.path.join(inc, "fmt", "format.h")) else None
    except Exception:
        return None


def available():
    return os.path.isdir(os.path.join(REF, "altro")) and fmt_include() is not None


def sources():
    out = []
    for pattern in ("altro/**/*.cpp", "examples/*.cpp", "examples/problems/*.cpp"):
        for path in sorted(glob.glob(os.path.join(REF, pattern), recursive=True)):
            if os.path.relpath(path, REF) not in SKIP:
                out.append(path)
    return out + [os.path.join(ROOT, "oracle", "ref_shim", "ref_driver.cpp")]


def build(force=False):
    """-> path of the library.  Raises CalledProcessError with the compiler output on failure."""
    srcs = sources()
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs + [__file__]):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    flags = ["g++", "-std=c++14", "-O2", "-fPIC", "-w", "-DNDEBUG", "-DFMT_HEADER_ONLY",
             "-include", os.path.join(ROOT, "oracle", "ref_shim", "compat.hpp"),
             "-I", REF, "-I", os.path.join(ROOT, "altro_cpp_b200", "host", "include"), "-I", fmt_include()]

    def compile_one(src):
        obj = os.path.join(OUT, os.path.relpath(src, "/").replace("/", "_") + ".o")
        r = subprocess.run(flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, src, output=r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(compile_one, srcs))
    r = subprocess.run(["g++", "-shared", "-o", LIB] + objs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise subprocess.CalledProcessError(r.returncode, "link", output=r.stdout + r.stderr)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    if not available():
        sys.exit("the reference sources (or fmt's headers) are not available here")
    try:
        print(build(force="--force" in sys.argv))
    except subprocess.CalledProcessError as e:
        print(e.output[-6000:])
        sys.exit(1)
