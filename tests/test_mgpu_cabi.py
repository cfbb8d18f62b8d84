"""kiss_fftnd over several GPUs through the C-ABI alone: tests/cpp/test_mgpu.c is a plain C program (one process per
GPU, rendezvous id passed through pipes) linked against libkissfft-float.so.  One rank runs on any GPU box; the
multi-rank cases need as many GPUs and are skipped otherwise (run them with `gpurun --gpus 2`)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "test_mgpu")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_exe():
    import kissfft_b200
    from kissfft_b200 import build as kbuild
    kbuild.build_one("float")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_mgpu.c")
    lib = kissfft_b200.lib_path("float")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["gcc", "-O2", "-std=gnu11", "-Dkiss_fft_scalar=float", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(CUDA_HOME, "include"), src, "-o", EXE, lib, "-L", os.path.join(CUDA_HOME, "lib64"),
                        "-lcudart", "-lm", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64")],
                       check=True)
    return EXE


def test_c_program_builds():
    """CPU: the C test program compiles and links against the library and the headers alone"""
    assert os.path.exists(build_exe())


def run(G, dims, p2p, iters=0):
    r = subprocess.run([build_exe(), str(G), *map(str, dims), str(int(p2p)), str(iters)], capture_output=True, text=True, timeout=600)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert len(lines) == G
    return lines


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(64, 32, 128), (128, 128, 128), (16, 1000, 64)])
def test_single_rank_through_c_abi(dims):
    out = run(1, dims, False)
    assert out[0]["rel_rms"] <= out[0]["tol"]


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("G,dims", [(2, (64, 32, 128)), (2, (256, 256, 256)), (4, (128, 64, 256)), (8, (256, 128, 256))])
def test_multi_rank_through_c_abi(G, dims, p2p):
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip("needs %d GPUs" % G)
    out = run(G, dims, p2p)
    assert all(o["rel_rms"] <= o["tol"] for o in out)
    if p2p:
        assert all(o["p2p"] == 1 for o in out), "peer mapping (CUDA IPC) was expected to work on one node"
