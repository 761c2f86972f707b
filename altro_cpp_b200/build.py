"""Builds the CUDA extension in-tree: altro_cpp_b200/libaltro_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaltro_b200.so")
SOURCES = ["altro_b200.cu", "altro_b200_large.cu", "altro_b200_large_mma.cu"]
# per-source extra flags: the large-state path is compiled without FMA contraction (bit parity)
EXTRA = {"altro_b200_large.cu": ["-fmad=false"]}


def _headers():
    """Every header a source can include: all of csrc/ plus the public C ABI (globbed, so a new header cannot be missed)."""
    hs = [f for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".hpp"))]
    return hs + [os.path.join("..", "..", "include", "altro_b200.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-split-compile", "0",  # the kernels are independent: let nvcc optimise / assemble them in parallel
    "-Xfatbin", "-compress-all",  # the cubins (with their -lineinfo tables) compress ~3x; decompressed at load time
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + _headers():
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    if out == LIB and not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    base = [f for f in NVCC_FLAGS if f != "-shared"]
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(HERE, "build_" + os.path.splitext(src)[0] + ("" if out == LIB else "_" + os.path.basename(out)) + ".o")
        cmd = [nvcc] + base + EXTRA.get(src, []) + (["-Xptxas", "-v"] if verbose else []) + \
              [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        if os.environ.get("ALTRO_B200_FMAD", "1") == "0" and "-fmad=false" not in cmd:
            cmd.insert(1, "-fmad=false")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", out] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc link failed building libaltro_b200.so")
    return out


DEV_LIB = os.path.join(HERE, "libaltro_b200_dev.so")


def build_dev(verbose: bool = False) -> str:
    """Kernel-tuning build: unicycle, tile width 8 only (ALTRO_DEV_BUILD).  Use with ALTRO_B200_LIB=<path>."""
    return build(force=True, verbose=verbose, out=DEV_LIB, defines=("ALTRO_DEV_BUILD",))


if __name__ == "__main__":
    if "--dev" in sys.argv:
        print(build_dev(verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
