"""CPU-side checks of the product library: it loads, exports every symbol include/*.h declares, and its C
planner (radix schedule, twiddle tables, cfg placement protocol) matches the oracle -- no GPU, no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

import kissfft_b200
from kissfft_b200 import build as kbuild
from oracle.loader import TYPES, Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    kbuild.build_all()


@pytest.fixture(scope="module", params=TYPES)
def ctx(request):
    return request.param, kissfft_b200.get(request.param), Oracle(request.param)


class CfgHeader(ctypes.Structure):
    _fields_ = [("magic", ctypes.c_uint32), ("nfft", ctypes.c_int), ("inverse", ctypes.c_int), ("nstages", ctypes.c_int),
                ("factors", ctypes.c_int * 64)]


def test_exports_every_declared_symbol(ctx):
    tname, lib, _ = ctx
    declared = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"KISS_FFT_API\s*\*?\s*((?:kiss|kfc)_\w+)\s*\(", text))
    assert len(declared) >= 40
    assert declared == set(kissfft_b200.API_SYMBOLS) | set(kissfft_b200.FASTCONV_SYMBOLS)
    for name in declared:
        if name in kissfft_b200.FASTCONV_SYMBOLS and tname.startswith("int"):
            assert not hasattr(lib.lib, name), name          # float / double builds only
        else:
            assert hasattr(lib.lib, name), name


def test_datatype_of_build(ctx):
    tname, lib, _ = ctx
    assert lib.lib.kiss_fft_cuda_scalar_bytes() == np.dtype(lib.dtype).itemsize
    assert lib.lib.kiss_fft_cuda_is_fixed_point() == (1 if tname.startswith("int") else 0)


@pytest.mark.parametrize("nfft", [1, 2, 4, 7, 16, 30, 74, 120, 1000, 1024, 1155, 1800, 2048, 4096, 1009])
def test_planner_matches_oracle(ctx, nfft):
    """factor order (kiss_fft.c:306-328) and twiddles (kiss_fft.c:361-367) are bit-identical to the oracle's"""
    tname, lib, o = ctx
    for inverse in (0, 1):
        cfg = lib.alloc(nfft, inverse)
        hdr = CfgHeader.from_address(cfg)
        assert (hdr.nfft, hdr.inverse) == (nfft, inverse)
        fac = o.factor(nfft)
        assert hdr.nstages == len(fac)
        assert [(hdr.factors[2 * s], hdr.factors[2 * s + 1]) for s in range(len(fac))] == fac
        tw = np.ctypeslib.as_array((ctypes.c_byte * (nfft * 2 * np.dtype(lib.dtype).itemsize)).from_address(
            cfg + ctypes.sizeof(CfgHeader))).view(lib.dtype).reshape(nfft, 2)
        assert np.array_equal(tw, o.twiddles(nfft, inverse))
        lib.free(cfg)


def test_alloc_protocol(ctx):
    """mem/lenmem placement protocol of kiss_fft.h:94-115 and its users"""
    tname, lib, _ = ctx
    L = lib.lib
    need = ctypes.c_size_t(0)
    assert L.kiss_fft_alloc(360, 0, None, ctypes.byref(need)) is None
    assert need.value >= ctypes.sizeof(CfgHeader) + 359 * 2 * np.dtype(lib.dtype).itemsize
    buf = ctypes.create_string_buffer(need.value + 16)
    small = ctypes.c_size_t(need.value - 1)
    assert L.kiss_fft_alloc(360, 0, buf, ctypes.byref(small)) is None and small.value == need.value
    ok = ctypes.c_size_t(need.value + 16)
    assert L.kiss_fft_alloc(360, 0, buf, ctypes.byref(ok)) == ctypes.addressof(buf) and ok.value == need.value
    assert L.kiss_fft_alloc(0, 0, None, None) is None
    # real: odd length is refused (kiss_fftr.c:29-32); size query works
    assert L.kiss_fftr_alloc(33, 0, None, None) is None
    needr = ctypes.c_size_t(0)
    assert L.kiss_fftr_alloc(64, 0, None, ctypes.byref(needr)) is None and needr.value > 0
    bufr = ctypes.create_string_buffer(needr.value)
    assert L.kiss_fftr_alloc(64, 0, bufr, ctypes.byref(needr)) == ctypes.addressof(bufr)
    # N-D
    dims = (ctypes.c_int * 3)(6, 10, 4)
    neednd = ctypes.c_size_t(0)
    assert L.kiss_fftnd_alloc(dims, 3, 0, None, ctypes.byref(neednd)) is None and neednd.value > 0
    cfg = L.kiss_fftnd_alloc(dims, 3, 0, None, None)
    assert cfg
    lib.free(cfg)
    dimsr = (ctypes.c_int * 2)(6, 9)
    assert L.kiss_fftndr_alloc(dimsr, 2, 0, None, None) is None      # odd real axis
    dimsr = (ctypes.c_int * 2)(6, 10)
    cfg = L.kiss_fftndr_alloc(dimsr, 2, 0, None, None)
    assert cfg
    lib.free(cfg)


def test_next_fast_size(ctx):
    _, lib, _ = ctx
    for n, want in [(1, 1), (7, 8), (11, 12), (1001, 1024), (1155, 1200), (4097, 4320)]:
        assert lib.next_fast_size(n) == want


def test_plan_table(ctx):
    """the benchmark lengths must be served by compile-time fused plans"""
    _, lib, _ = ctx
    for n in (1024, 2048, 1000, 1155, 256, 64):
        assert lib.lib.kiss_fft_cuda_plan_kind(n) == 1


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(kissfft_b200, "HERE", str(tmp_path))
    with pytest.raises(kissfft_b200.KissFFTError):
        kissfft_b200.KissFFT("float")


@pytest.mark.parametrize("howmany,rows", [(32768, 1984), (65536, 4096), (100000, 3584), (7, 64), (1, 1), (4096, 512), (11905, 1984), (3071, 512)])
@pytest.mark.parametrize("ramp", [0, 1])
def test_host_pipeline_chunks_cover_the_batch(howmany, rows, ramp):
    """kf_host_pipeline's chunk list (ramped at both ends so that the PCIe fill/drain of a call is short): every row
    exactly once, in order, no chunk larger than the staging buffers"""
    import ctypes
    import kissfft_b200
    L = kissfft_b200.get("float").lib
    L.kiss_fft_cuda_debug_chunks.restype = ctypes.c_size_t
    L.kiss_fft_cuda_debug_chunks.argtypes = [ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                             ctypes.POINTER(ctypes.c_size_t), ctypes.c_size_t]
    cap = 4096
    first, n = (ctypes.c_size_t * cap)(), (ctypes.c_size_t * cap)()
    k = L.kiss_fft_cuda_debug_chunks(howmany, rows, ramp, first, n, cap)
    assert 0 < k <= cap
    pos = 0
    for i in range(k):
        assert first[i] == pos and 0 < n[i] <= rows
        pos += n[i]
    assert pos == howmany
    if ramp and howmany >= 6 * rows and rows >= 512:
        assert n[0] < rows / 4 and n[k - 1] < rows / 4, "the first and the last chunk are the exposed ones"


@pytest.mark.parametrize("cols,nranks,want,tail16", [(512, 2, 2, 4), (128, 8, 2, 4), (256, 4, 2, 4), (128, 8, 4, 4), (64, 2, 2, 4), (32, 2, 2, 4),
                                                     (16, 2, 2, 4), (125, 8, 2, 4), (512, 2, 2, 0), (512, 1, 4, 4), (384, 2, 8, 2)])
def test_mgpu_column_chunks(cols, nranks, want, tail16):
    """chunks of the k2 columns of kiss_fftnd_mgpu_exec: cover [0, cols) in order, whole 16-column tiles when there is more than
    one chunk, and a narrower last chunk (the one whose axis-0 pass nothing overlaps) when asked for"""
    import ctypes
    import kissfft_b200
    L = kissfft_b200.get("float").lib
    coff = (ctypes.c_int * 16)()
    n = L.kiss_fftnd_mgpu_debug_chunks(cols, nranks, want, tail16, coff, 16)
    assert 1 <= n <= max(1, want)
    assert coff[0] == 0 and coff[n] == cols
    w = [coff[j + 1] - coff[j] for j in range(n)]
    assert all(x > 0 for x in w)
    if nranks == 1:
        assert n == 1
    if n > 1:
        assert all(x % 16 == 0 for x in w)
        if tail16:
            assert w[-1] <= min(w[:-1])
