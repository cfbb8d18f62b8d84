"""Index-math tests of the kernel bodies WITHOUT a GPU: the template code of kissfft_b200/csrc/kf_body.h is
executed by tests/emul (one std::thread per CUDA thread) and compared with the oracle.  This proves the
addressing / plan logic, not the GPU build -- the -m gpu tests compare the real kernels with the oracle."""
import numpy as np
import pytest

from oracle.loader import TYPES, Oracle, random_input, rel_rms
from tests.emul_util import C2C, C2C_COL, C2R, R2C, Emulator

TOL = {"float": 1e-6, "double": 1e-14}


def check(tname, got, want, nfft):
    if tname in TOL:
        assert rel_rms(got, want) <= TOL[tname] * max(1.0, np.log2(nfft))
    else:
        assert np.array_equal(got, want)


@pytest.fixture(scope="module", autouse=True)
def _emulators_built():
    from tests.emul_util import build_all
    build_all()


@pytest.fixture(scope="module", params=TYPES)
def env(request):
    return request.param, Oracle(request.param), Emulator(request.param)


def test_fused_c2c(env):
    tname, o, em = env
    for nfft, modes in em.plans():
        if C2C not in modes:
            continue
        for inverse in (0, 1):
            howmany = 11
            x = random_input(tname, (howmany, nfft), 100 + nfft)
            out = np.zeros_like(x)
            em.fused(nfft, C2C, inverse, x, out, howmany, nfft, nfft, 1, o.twiddles(nfft, inverse), factors=o.factor(nfft))
            check(tname, out, o.fft(x, inverse), nfft)


def test_generic_c2c_strided_input(env):
    """kiss_fft_stride with in_stride != 1 is served by the run-time kernel (the fused C2C plans assume contiguous rows)"""
    tname, o, em = env
    nfft, howmany, stride = 64, 5, 3
    x = random_input(tname, (howmany, nfft * stride), 7)
    out = np.zeros((howmany, nfft, 2), x.dtype)
    em.generic(nfft, C2C, 0, o.factor(nfft), x, out, howmany, nfft * stride, nfft, stride, o.twiddles(nfft, 0))
    check(tname, out, o.fft(x, 0, in_stride=stride, nfft=nfft), nfft)


def test_fused_columns(env):
    tname, o, em = env
    for nfft, modes in em.plans():
        if C2C_COL not in modes or nfft > 1024:
            continue
        ncols = 37
        x = random_input(tname, (nfft, ncols), 200 + nfft)           # [nfft][ncols]: column i strided by ncols
        out = np.zeros((ncols, nfft, 2), x.dtype)
        em.fused(nfft, C2C_COL, 0, x, out, ncols, 1, nfft, ncols, o.twiddles(nfft, 0))
        want = o.fft(np.ascontiguousarray(x.transpose(1, 0, 2)), 0)
        check(tname, out, want, nfft)


def test_fused_real(env):
    tname, o, em = env
    for nc, modes in em.plans():
        if R2C not in modes:
            continue
        nfft, howmany = 2 * nc, 9
        x = random_input(tname, (howmany, nfft), 300 + nc, complex_=False)
        X = np.zeros((howmany, nc + 1, 2), x.dtype)
        em.fused(nc, R2C, 0, x, X, howmany, nc, nc + 1, 1, o.twiddles(nc, 0), o.super_twiddles(nc, 0), factors=o.factor(nc))
        want = o.fftr(x)
        check(tname, X, want, nfft)
        spec = want if tname in TOL else random_input(tname, (howmany, nc + 1), 301 + nc)
        y = np.zeros((howmany, nfft), x.dtype)
        em.fused(nc, C2R, 1, spec, y, howmany, nc + 1, nc, 1, o.twiddles(nc, 1), o.super_twiddles(nc, 1), factors=o.factor(nc))
        check(tname, y, o.fftri(spec), nfft)


@pytest.mark.parametrize("n1,n2", [(128, 128), (256, 128)])
def test_fourstep_long_rows(env, n1, n2):
    """N = n1*n2 as two fused column passes (float / double): columns of length n1 times W_N^(n2 k1), then columns of
    length n2 written in natural order"""
    tname, o, em = env
    if tname not in TOL:
        pytest.skip("the four-step path is float / double only (a different factorisation is not bit-exact)")
    N, rows = n1 * n2, 2
    for inverse in (0, 1):
        x = random_input(tname, (rows, N), 77 + inverse)
        out = np.zeros_like(x)
        ok = em.fourstep(n1, n2, inverse, x, out, rows, o.twiddles(n1, inverse), o.twiddles(n2, inverse), o.twiddles(N, inverse))
        assert ok, "no fused column plan for %d or %d" % (n1, n2)
        check(tname, out, o.fft(x, inverse), N)


def test_fftnd_in_layout(env):
    """kiss_fftnd with every axis transformed where it lies (kf_api.c:kf_fftnd_dev_inlayout): axes 0 and 1 by the
    column-in / column-out pass, the last axis as rows; same bits as the reference's transposing sweeps in fixed point"""
    tname, o, em = env
    dims = (64, 128, 4)
    x = random_input(tname, dims, 31)
    a = np.zeros_like(x)
    assert em.colcol(64, 0, x, a, 1, 128 * 4, o.twiddles(64, 0))                 # axis 0: one plane of 512 columns
    b = np.zeros_like(x)
    assert em.colcol(128, 0, a, b, 64, 4, o.twiddles(128, 0))                    # axis 1: 64 planes of 4 columns (ragged tile)
    c = np.zeros_like(x)
    em.generic(4, C2C, 0, o.factor(4), b.reshape(-1, 4, 2), c.reshape(-1, 4, 2), 64 * 128, 4, 4, 1, o.twiddles(4, 0))   # axis 2: rows
    check(tname, c, o.fftnd(x), 64 * 128 * 4)


def test_column_ring_plans(env):
    """the tensor-map (TMA) input ring of the column modes: tiles of adjacent columns land as [row][column] boxes; the
    transposing pass (kiss_fftnd.c:172-178) and the layout-keeping pass must match the oracle's 1-D transforms"""
    tname, o, em = env
    nfft, nplanes, ncols = 1024, 2, 24
    x = random_input(tname, (nplanes, nfft, ncols), 77)
    want = o.fft(np.ascontiguousarray(x.transpose(0, 2, 1, 3)).reshape(nplanes * ncols, nfft, 2)).reshape(nplanes, ncols, nfft, 2)
    out = np.zeros((nplanes, ncols, nfft, 2), x.dtype)
    rc = em.colring(nfft, 1, 0, x, out, nplanes, ncols, ncols, nfft * ncols, ncols * nfft, nfft, o.twiddles(nfft, 0), nblocks=2)
    if rc == -1:
        pytest.skip("no column-ring plan for this datatype build")
    assert rc >= 0
    check(tname, out, want, nfft)
    out2 = np.zeros_like(x)
    rc = em.colring(nfft, 5, 0, x, out2, nplanes, ncols, ncols, nfft * ncols, nfft * ncols, 1, o.twiddles(nfft, 0), nblocks=3)
    assert rc >= 0
    check(tname, out2, want.transpose(0, 2, 1, 3), nfft)


def test_experimental_real_plans(env):
    """plan variants that are not in the product list yet (tests/emul/experimental_plans.h): paired groups with the
    even/odd lane mapping, with and without the input stage as second exchange buffer"""
    tname, o, em = env
    ran = 0
    for idx, nc, modes in em.experimental():
        nfft, howmany = 2 * nc, 7
        if C2C in modes:
            for inverse in (0, 1):
                x = random_input(tname, (howmany, nc), 903 + idx)
                out = np.zeros_like(x)
                em.fused(nc, C2C, inverse, x, out, howmany, nc, nc, 1, o.twiddles(nc, inverse), factors=o.factor(nc), experimental=idx)
                check(tname, out, o.fft(x, inverse), nc)
                ran += 1
        if R2C in modes:
            x = random_input(tname, (howmany, nfft), 900 + idx, complex_=False)
            X = np.zeros((howmany, nc + 1, 2), x.dtype)
            em.fused(nc, R2C, 0, x, X, howmany, nc, nc + 1, 1, o.twiddles(nc, 0), o.super_twiddles(nc, 0), factors=o.factor(nc),
                     experimental=idx)
            check(tname, X, o.fftr(x), nfft)
            ran += 1
        if C2R in modes:
            spec = o.fftr(random_input(tname, (howmany, nfft), 901 + idx, complex_=False)) if tname in TOL \
                else random_input(tname, (howmany, nc + 1), 902 + idx)
            y = np.zeros((howmany, nfft), spec.dtype)
            em.fused(nc, C2R, 1, spec, y, howmany, nc + 1, nc, 1, o.twiddles(nc, 1), o.super_twiddles(nc, 1), factors=o.factor(nc),
                     experimental=idx)
            check(tname, y, o.fftri(spec), nfft)
            ran += 1
    assert ran >= 2


@pytest.mark.parametrize("nfft", [1, 2, 3, 5, 7, 12, 30, 74, 120, 143, 360])
def test_generic_c2c(env, nfft):
    tname, o, em = env
    for inverse in (0, 1):
        howmany = 5
        x = random_input(tname, (howmany, nfft), 400 + nfft)
        out = np.zeros_like(x)
        em.generic(nfft, C2C, inverse, o.factor(nfft), x, out, howmany, nfft, nfft, 1, o.twiddles(nfft, inverse))
        check(tname, out, o.fft(x, inverse), nfft)


def test_generic_columns_and_real(env):
    tname, o, em = env
    nfft, ncols = 30, 7
    x = random_input(tname, (nfft, ncols), 9)
    out = np.zeros((ncols, nfft, 2), x.dtype)
    em.generic(nfft, C2C_COL, 0, o.factor(nfft), x, out, ncols, 1, nfft, ncols, o.twiddles(nfft, 0), tpc=3)
    check(tname, out, o.fft(np.ascontiguousarray(x.transpose(1, 0, 2)), 0), nfft)
    for n in (2, 4, 6, 30, 120):
        nc, howmany = n // 2, 5
        xr = random_input(tname, (howmany, n), 500 + n, complex_=False)
        X = np.zeros((howmany, nc + 1, 2), xr.dtype)
        em.generic(nc, R2C, 0, o.factor(nc), xr, X, howmany, nc, nc + 1, 1, o.twiddles(nc, 0), o.super_twiddles(nc, 0))
        want = o.fftr(xr)
        check(tname, X, want, n)
        spec = want if tname in TOL else random_input(tname, (howmany, nc + 1), 501 + n)
        y = np.zeros((howmany, n), xr.dtype)
        em.generic(nc, C2R, 1, o.factor(nc), spec, y, howmany, nc + 1, nc, 1, o.twiddles(nc, 1), o.super_twiddles(nc, 1))
        check(tname, y, o.fftri(spec), n)


def test_q15_full_scale_wraparound():
    """Q15 inputs at full scale (including -32768) make the reference's int16 stores wrap; the kernels re-create the
    truncation exactly where a wrapped value feeds a multiply or a shift, so they must still agree bit for bit."""
    tname = "int16_t"
    o, em = Oracle(tname), Emulator(tname)
    rng = np.random.default_rng(5)
    for nfft in (64, 256, 1000, 1024, 1155, 2048):
        howmany = 4
        x = rng.integers(-32768, 32768, size=(howmany, nfft, 2), dtype=np.int64).astype(np.int16)
        x[0, :8] = -32768
        x[1, :8] = 32767
        out = np.zeros_like(x)
        em.fused(nfft, C2C, 0, x, out, howmany, nfft, nfft, 1, o.twiddles(nfft, 0), factors=o.factor(nfft))
        assert np.array_equal(out, o.fft(x, 0)), nfft
        out2 = np.zeros_like(x)
        em.generic(nfft, C2C, 0, o.factor(nfft), x, out2, howmany, nfft, nfft, 1, o.twiddles(nfft, 0))
        assert np.array_equal(out2, o.fft(x, 0)), nfft
    for n in (128, 2000, 4096):
        nc = n // 2
        xr = rng.integers(-32768, 32768, size=(3, n), dtype=np.int64).astype(np.int16)
        X = np.zeros((3, nc + 1, 2), np.int16)
        em.fused(nc, R2C, 0, xr, X, 3, nc, nc + 1, 1, o.twiddles(nc, 0), o.super_twiddles(nc, 0), factors=o.factor(nc))
        assert np.array_equal(X, o.fftr(xr)), n
        S = rng.integers(-32768, 32768, size=(3, nc + 1, 2), dtype=np.int64).astype(np.int16)
        y = np.zeros((3, n), np.int16)
        em.fused(nc, C2R, 1, S, y, 3, nc + 1, nc, 1, o.twiddles(nc, 1), o.super_twiddles(nc, 1), factors=o.factor(nc))
        assert np.array_equal(y, o.fftri(S)), n


@pytest.mark.parametrize("nfft", [2, 7, 12, 30, 143, 360, 1000, 1024, 1155])
def test_multipass_stages(env, nfft):
    """the global-memory multi-pass path (one radix stage per launch) used for lengths beyond shared memory"""
    tname, o, em = env
    howmany, stride = 3, 2
    for inverse in (0, 1):
        x = random_input(tname, (howmany, nfft * stride + 1), 600 + nfft)
        out = np.zeros((howmany, nfft + 2, 2), x.dtype)
        em.multipass(nfft, inverse, o.factor(nfft), x, out, howmany, nfft * stride + 1, nfft + 2, stride, o.twiddles(nfft, inverse))
        check(tname, out[:, :nfft], o.fft(x[:, : nfft * stride], inverse, in_stride=stride, nfft=nfft), nfft)
        assert not out[:, nfft:].any()


def test_multipass_real(env):
    tname, o, em = env
    for n in (4, 30, 1000):
        nc, howmany = n // 2, 3
        xr = random_input(tname, (howmany, n), 700 + n, complex_=False)
        T = np.zeros((howmany, nc, 2), xr.dtype)
        em.multipass(nc, 0, o.factor(nc), xr, T, howmany, nc, nc, 1, o.twiddles(nc, 0))
        X = np.zeros((howmany, nc + 1, 2), xr.dtype)
        em.realpass(nc, 1, T, X, howmany, nc, nc + 1, o.super_twiddles(nc, 0))
        want = o.fftr(xr)
        check(tname, X, want, n)
        spec = want if tname in TOL else random_input(tname, (howmany, nc + 1), 701 + n)
        T2 = np.zeros((howmany, nc, 2), xr.dtype)
        em.realpass(nc, 0, spec, T2, howmany, nc + 1, nc, o.super_twiddles(nc, 1))
        y = np.zeros((howmany, n), xr.dtype)
        em.multipass(nc, 1, o.factor(nc), T2, y, howmany, nc, nc, 1, o.twiddles(nc, 1))
        check(tname, y, o.fftri(spec), n)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft,nimp", [(256, 33), (512, 100), (1024, 129), (2048, 513), (4096, 1000)])
def test_fused_fast_convolution_emulated(tname, nfft, nimp):
    """fastconv_body (FFT -> .*H -> IFFT in one kernel, spectrum handed over in registers) vs the oracle pipeline"""
    o, em = Oracle(tname), Emulator(tname)
    rng = np.random.default_rng(3)
    imp = rng.uniform(-1, 1, size=(nimp, 2)).astype(o.dtype) / nimp
    ngood = nfft - nimp + 1
    nblocks = 9
    x = random_input(tname, ((nblocks - 1) * ngood + nfft,), 17)
    rot = np.zeros((nfft, 2), o.dtype)
    rot[0] = imp[nimp - 1]
    rot[nfft - nimp + 1:] = imp[: nimp - 1]
    H = (o.fft(rot) * np.float32(1.0 / nfft)).astype(o.dtype)
    out = np.zeros_like(x)
    em.fastconv(nfft, x, out, nblocks, ngood, H, o.twiddles(nfft, 0), o.twiddles(nfft, 1))
    blocks = np.stack([x[b * ngood: b * ngood + nfft] for b in range(nblocks)])
    X = o.fft(blocks).astype(np.float64)
    Hc = H.astype(np.float64)
    Y = np.empty_like(X)
    Y[..., 0] = X[..., 0] * Hc[:, 0] - X[..., 1] * Hc[:, 1]
    Y[..., 1] = X[..., 0] * Hc[:, 1] + X[..., 1] * Hc[:, 0]
    y = o.fft(Y.astype(o.dtype), True)[:, :ngood].reshape(-1, 2)
    assert rel_rms(out[: nblocks * ngood], y) <= 3 * TOL[tname] * np.log2(nfft)
    assert not out[nblocks * ngood:].any()
