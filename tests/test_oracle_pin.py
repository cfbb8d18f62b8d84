"""Pins the oracle (oracle/kiss_oracle.c): (1) against the golden vectors generated from the compiled,
unmodified reference and committed under tests/golden/, everywhere; (2) against the compiled reference itself
(oracle/_ref) on fresh random inputs wherever it is available (the build container; the GPU box gets the
prebuilt .so with the snapshot).  Bit-exact in all four datatypes -- the restatement performs the same
floating-point operations in the same order, so even float/double agree to the last bit."""
import numpy as np
import pytest

from oracle.loader import TYPES, Oracle, Reference, have_reference, random_input
from tests import golden_util

FLOATS = ("float", "double")


@pytest.fixture(scope="module", params=TYPES)
def tname(request):
    return request.param


def test_golden_c2c(tname):
    o = Oracle(tname)
    n = 0
    for key, x, y in golden_util.cases(tname, "c2c"):
        nfft, inv = map(int, key.split("_"))
        assert np.array_equal(o.fft(x, inv), y), key
        n += 1
    assert n >= 10


def test_golden_real(tname):
    o = Oracle(tname)
    for key, x, y in golden_util.cases(tname, "r2c"):
        assert np.array_equal(o.fftr(x), y), key
    for key, x, y in golden_util.cases(tname, "c2r"):
        assert np.array_equal(o.fftri(x), y), key


def test_golden_nd(tname):
    o = Oracle(tname)
    for key, x, y in golden_util.cases(tname, "nd"):
        inv = int(key.split("_")[-1])
        assert np.array_equal(o.fftnd(x, inv), y), key
    for key, x, y in golden_util.cases(tname, "ndr"):
        assert np.array_equal(o.fftndr(x), y), key
    for key, x, y in golden_util.cases(tname, "ndri"):
        assert np.array_equal(o.fftndri(x), y), key


needs_ref = pytest.mark.skipif(not have_reference(), reason="compiled reference (oracle/_ref) not present")


@needs_ref
def test_vs_compiled_reference_1d(tname):
    o, r = Oracle(tname), Reference(tname)
    for n in [1, 2, 3, 4, 5, 7, 16, 30, 74, 120, 143, 148, 1000, 1024, 1155, 1800, 2048]:
        for inv in (0, 1):
            x = random_input(tname, (3, n), 10 + n)
            assert np.array_equal(o.fft(x, inv), r.fft(x, inv)), (n, inv)
    x = random_input(tname, (7 * 5 * 3,), 5)
    assert np.array_equal(o.fft(x, 0, in_stride=5, nfft=21), r.fft_stride(x, 21, 5))


@needs_ref
def test_vs_compiled_reference_real_and_nd(tname):
    o, r = Oracle(tname), Reference(tname)
    for n in [2, 4, 6, 30, 120, 1000, 4096, 2310]:
        x = random_input(tname, (2, n), n, complex_=False)
        a = o.fftr(x)
        assert np.array_equal(a, r.fftr(x)), n
        S = a if tname in FLOATS else random_input(tname, (2, n // 2 + 1), n + 1)
        assert np.array_equal(o.fftri(S), r.fftri(S)), n
    for dims in [(4, 3), (2, 3, 4), (30, 20, 12), (16, 16, 16), (8,), (5, 6, 7, 4)]:
        for inv in (0, 1):
            x = random_input(tname, dims, 3)
            assert np.array_equal(o.fftnd(x, inv), r.fftnd(x, inv)), dims
    for dims in [(4, 6), (2, 3, 4), (30, 20, 12), (5, 6, 8)]:
        x = random_input(tname, dims, 3, complex_=False)
        a = o.fftndr(x)
        assert np.array_equal(a, r.fftndr(x)), dims
        S = a if tname in FLOATS else random_input(tname, a.shape[:-1], 9)
        assert np.array_equal(o.fftndri(S), r.fftndri(S)), dims


def test_reference_style_checks(tname):
    """the reference's own test ideas at our sizes: vs numpy SNR (test/testkiss.py:24-31,85-89) and
    kiss_fftr vs kiss_fft of the same real data (test/test_real.c:101-117)"""
    o = Oracle(tname)
    n = 1024
    x = random_input(tname, (n,), 2)
    scale = 1.0 if tname in FLOATS else 1.0 / n
    want = np.fft.fft(x[:, 0].astype(np.float64) + 1j * x[:, 1]) * scale
    got = o.fft(x).astype(np.float64)
    err = (got[:, 0] + 1j * got[:, 1]) - want
    snr = 10 * np.log10(np.sum(np.abs(want) ** 2) / max(np.sum(np.abs(err) ** 2), 1e-300))
    assert snr >= (10 if tname == "int16_t" else 90)
    xr = random_input(tname, (n,), 3, complex_=False)
    as_cpx = np.stack([xr, np.zeros_like(xr)], -1)
    a = o.fftr(xr).astype(np.float64)
    b = o.fft(as_cpx).astype(np.float64)[: n // 2 + 1]
    if tname in FLOATS:
        assert np.allclose(a, b, atol=1e-3 if tname == "float" else 1e-10)
    else:
        assert np.max(np.abs(a - b)) <= 4          # different rounding paths, a few LSB
