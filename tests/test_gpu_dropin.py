"""Drop-in proof on the GPU: the reference's OWN, UNMODIFIED programs -- tools/fftutil.c (the `fft-<type>` CLI that the
reference's test/testkiss.py drives), test/test_real.c, test/twotonetest.c, test/benchkiss.c -- compiled by
`make -C oracle dropin` against this repo's headers and linked against this repo's CUDA libraries, run as black boxes.
The binaries live under oracle/_ref/dropin (git-ignored, shipped with the snapshot); where they are absent the tests skip."""
import os
import subprocess

import numpy as np
import pytest

from oracle.loader import TYPES, Oracle, dropin_path, random_input

pytestmark = pytest.mark.gpu


def _need(name, tname):
    p = dropin_path(name, tname)
    if not os.path.exists(p):
        pytest.skip("%s not built (needs /root/reference at build time)" % p)
    return p


@pytest.fixture(params=TYPES)
def tname(request):
    return request.param


def _pipe(exe, args, data):
    r = subprocess.run([exe, *args], input=data.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-500:]
    return r.stdout


def check(tname, got, want):
    if tname in ("float", "double"):
        err = np.sqrt(np.sum((got.astype(np.float64) - want) ** 2) / np.sum(want.astype(np.float64) ** 2))
        assert err <= (1e-6 if tname == "float" else 1e-14) * 10
    else:
        assert np.array_equal(got, want)


def test_fftutil_cli(tname):
    """raw stdin -> stdout like the reference's test/testkiss.py drives it: 1-D, N-D, inverse, real (-R)"""
    exe = _need("fft", tname)
    o = Oracle(tname)
    x = random_input(tname, (3, 120), 1)                                   # three 120-point rows streamed through
    out = np.frombuffer(_pipe(exe, ["-n", "120"], x), o.dtype).reshape(3, 120, 2)
    check(tname, out, o.fft(x))
    out = np.frombuffer(_pipe(exe, ["-n", "120", "-i"], x), o.dtype).reshape(3, 120, 2)
    check(tname, out, o.fft(x, True))
    for dims in ((4, 6), (3, 4, 5), (16, 8, 4)):
        xn = random_input(tname, dims, 2)
        out = np.frombuffer(_pipe(exe, ["-n", ",".join(map(str, dims))], xn), o.dtype).reshape(dims + (2,))
        check(tname, out, o.fftnd(xn))                                    # fftutil transforms in place (fftutil.c:51)
    xr = random_input(tname, (2, 240), 3, complex_=False)
    out = np.frombuffer(_pipe(exe, ["-n", "240", "-R"], xr), o.dtype).reshape(2, 121, 2)
    check(tname, out, o.fftr(xr))
    xr2 = random_input(tname, (6, 10, 8), 4, complex_=False)
    out = np.frombuffer(_pipe(exe, ["-n", "6,10,8", "-R"], xr2), o.dtype).reshape(6, 10, 5, 2)
    check(tname, out, o.fftndr(xr2))


def test_reference_test_real(tname):
    """test/test_real.c: kiss_fftr vs kiss_fft and kiss_fftri vs inverse kiss_fft; exits 1 when an SNR is below 10 dB"""
    exe = _need("test_real", tname)
    r = subprocess.run([exe], capture_output=True, timeout=600)
    assert r.returncode == 0, (r.stdout.decode()[-400:], r.stderr.decode()[-400:])
    assert b"snr" in r.stdout.lower()


def test_reference_twotone_and_bench(tname):
    exe = _need("twotone", tname)
    r = subprocess.run([exe], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-400:]
    bench = _need("benchkiss", tname)
    for args in (["-n", "1800", "-x", "200"], ["-n", "1024", "-x", "200", "-r"], ["-n", "32,32", "-x", "20"]):
        r = subprocess.run([bench, *args], capture_output=True, timeout=600)
        assert r.returncode == 0, (args, r.stderr.decode()[-400:])


@pytest.mark.parametrize("tool", ["fastconv", "fastconvr"])
def test_reference_fast_fir_tool(tname, tool, tmp_path):
    """tools/kiss_fastfir.c (overlap-scrap fast FIR filter) built against this repo's library must produce what the same
    program produces with the compiled reference library (oracle/_ref/refbin): bit-identical in Q15/Q31."""
    exe = _need(tool, tname)
    ref = os.path.join(os.path.dirname(os.path.dirname(exe)), "refbin", "%s-%s" % (tool, tname))
    if not os.path.exists(ref):
        pytest.skip("%s not built" % ref)
    cplx = tool == "fastconv"
    o = Oracle(tname)
    n, nh = 50000, 65
    x = random_input(tname, (n,), 11, complex_=cplx)
    h = random_input(tname, (nh,), 12, complex_=cplx)
    if tname in ("float", "double"):
        h = (h * 0.05).astype(o.dtype)
    else:
        h = (h // 8).astype(o.dtype)              # keep the Q15/Q31 filter gain below one
    fin, fh = tmp_path / "in.bin", tmp_path / "h.bin"
    x.tofile(fin)
    h.tofile(fh)
    outs = []
    for prog in (exe, ref):
        fo = tmp_path / ("out_%d.bin" % len(outs))
        r = subprocess.run([prog, "-i", str(fin), "-o", str(fo), "-h", str(fh)], capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-500:]
        outs.append(np.fromfile(fo, o.dtype))
    got, want = outs
    assert got.size == want.size == (n - nh + 1) * (2 if cplx else 1)
    check(tname, got, want)
