"""kiss_fftnd over several GPUs through the C-ABI alone: tests/cpp/test_mgpu.c is a plain C program (one process per
GPU, rendezvous id passed through pipes) linked against libkissfft-float.so.  One rank runs on any GPU box; the
multi-rank cases need as many GPUs and are skipped otherwise (run them with `gpurun --gpus 2`)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
P2P, REFERENCE_ORDER = 1, 2          # KISS_FFT_MGPU_* flags (include/kiss_fft_cuda.h)


def build_exe(tname="float"):
    import kissfft_b200
    from kissfft_b200 import build as kbuild
    kbuild.build_one(tname)
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_mgpu.c")
    lib = kissfft_b200.lib_path(tname)
    exe = os.path.join(BUILD, "test_mgpu" + ("" if tname == "float" else "-" + tname))
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["gcc", "-O2", "-std=gnu11", *[f for f in kbuild.TYPEFLAGS[tname] if f != "-DKF_IS_DOUBLE"],
                        "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(CUDA_HOME, "include"), src, "-o", exe, lib, "-L", os.path.join(CUDA_HOME, "lib64"),
                        "-lcudart", "-lm", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64")],
                       check=True)
    return exe


@pytest.mark.parametrize("tname", ["float", "int16_t"])
def test_c_program_builds(tname):
    """CPU: the C test program compiles and links against the library and the headers alone"""
    assert os.path.exists(build_exe(tname))


def run(G, dims, flags, iters=0, tname="float"):
    r = subprocess.run([build_exe(tname), str(G), *map(str, dims), str(int(flags)), str(iters)], capture_output=True, text=True, timeout=600)
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


@pytest.mark.gpu
@pytest.mark.parametrize("tname", ["float", "double", "int16_t", "int32_t"])
@pytest.mark.parametrize("dims", [(64, 32, 128), (16, 1000, 64), (1024, 16, 32)])
def test_reference_order_single_rank(tname, dims):
    """the exact mode (axes 0, 1, exchange, 2) on one rank: bit-identical to kiss_fftnd_dev in fixed point"""
    out = run(1, dims, REFERENCE_ORDER, tname=tname)
    assert out[0]["reference_order"] == 1 and out[0]["rel_rms"] <= out[0]["tol"]
    if tname.startswith("int"):
        assert out[0]["mismatches"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("tname", ["int16_t", "float"])
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("G,dims", [(2, (64, 32, 128)), (2, (256, 128, 256)), (4, (128, 64, 256))])
def test_reference_order_multi_rank(G, dims, p2p, tname):
    """SURVEY 8e "fixed-point N-D": Q15 across GPUs, bit-identical to the single-GPU kiss_fftnd (kiss_fftnd.c:156-188)"""
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip("needs %d GPUs" % G)
    out = run(G, dims, REFERENCE_ORDER | (P2P if p2p else 0), tname=tname)
    assert all(o["rel_rms"] <= o["tol"] for o in out)
    if tname.startswith("int"):
        assert all(o["mismatches"] == 0 for o in out)


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("G,dims", [(1, (256, 1, 512)), (1, (96, 1, 1000)), (2, (256, 1, 512)), (2, (1024, 1, 2048)), (2, (96, 1, 1000)), (4, (512, 1, 256))])
def test_2d_through_c_abi(G, dims, p2p):
    """north_star "large 2-D/3-D kiss_fftnd": ndims = 2 (d1 == 1 selects it in the test program) -- rows, transposing exchange,
    rows; compared with the single-GPU kiss_fftnd_dev of the whole array"""
    import torch
    if torch.cuda.device_count() < G:
        pytest.skip("needs %d GPUs" % G)
    out = run(G, dims, p2p)
    assert all(o["rel_rms"] <= o["tol"] for o in out)
