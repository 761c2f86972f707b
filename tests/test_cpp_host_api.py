"""The C-ABI used from host C++ (the reference's language) through include/altro_b200.hpp."""
import os
import subprocess

import pytest

import altro_cpp_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_smoke(tmp_path):
    pkg.lib()  # makes sure libaltro_b200.so exists
    exe = str(tmp_path / "host_api_smoke")
    libdir = os.path.join(ROOT, "altro_cpp_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_smoke.cpp"), "-o", exe,
                           "-L", libdir, "-laltro_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_host_api_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([build_smoke(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_host_api_solves_the_reference_problem(tmp_path):
    r = subprocess.run([build_smoke(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "iterations total/outer 50/5" in r.stdout
