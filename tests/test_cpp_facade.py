"""include/kissfft.hh, the C++ facade with the reference class's interface (reference kissfft.hh:16-189), compiled from
tests/cpp/test_kissfft_hh.cpp -- the counterpart of the reference's test/testcpp.cc."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_kissfft_hh.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_kissfft_hh")


def _build():
    import kissfft_b200.build as b
    b.build_all(("float", "double"))
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-std=c++11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-ldl", "-o", BIN], check=True)
    return dict(os.environ, KISSFFT_B200_LIB_DIR=os.path.join(ROOT, "kissfft_b200", "lib"))


def test_facade_compiles_and_has_no_cpu_path():
    """without a CUDA device the facade must throw, never compute on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    env = _build()
    r = subprocess.run([BIN, "--no-gpu"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("threw as expected") == 2


@pytest.mark.gpu
def test_facade_against_direct_dft():
    env = _build()
    r = subprocess.run([BIN], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
