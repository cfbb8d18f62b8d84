"""GPU parity tests: the CUDA path, called through the C-ABI of libkissfft-<type>.so, against the oracle
(oracle/kiss_oracle.c, itself pinned bit-for-bit to the compiled reference by tests/test_oracle_pin.py).

Bars (BASELINE.json north_star): Q15/Q31 bit-exact; float relative RMS <= 1e-6*log2(N); double <= 1e-14*log2(N).
"""
import ctypes
import os

import numpy as np
import pytest

from oracle.loader import TYPES, Oracle, Reference, have_reference, random_input, rel_rms

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

TOL = {"float": 1e-6, "double": 1e-14}


def check(tname, got, want, n, what=""):
    if tname in TOL:
        err = rel_rms(got, want)
        assert err <= TOL[tname] * max(1.0, np.log2(max(n, 2))), "%s %s n=%d rel-rms %.3g" % (tname, what, n, err)
    else:
        got = np.asarray(got)
        want = np.asarray(want)
        nbad = int(np.count_nonzero(got != want))
        assert nbad == 0, "%s %s n=%d: %d of %d scalars differ" % (tname, what, n, nbad, got.size)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


@pytest.fixture(scope="module", params=TYPES)
def ctx(request):
    import kissfft_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib = kissfft_b200.get(request.param)
    yield request.param, lib, Oracle(request.param)
    lib.force_generic(False)


C2C_SIZES = [1, 2, 3, 4, 5, 7, 8, 16, 30, 64, 74, 120, 143, 148, 256, 360, 1000, 1009, 1024, 1155, 1800, 2048, 4096]


@pytest.mark.parametrize("nfft", C2C_SIZES)
def test_c2c_batch(ctx, nfft):
    tname, lib, o = ctx
    howmany = 37
    for inverse in (False, True):
        x = random_input(tname, (howmany, nfft), 1000 + nfft)
        d_in, d_out = dev(x), dev(np.zeros_like(x))
        cfg = lib.alloc(nfft, inverse)
        lib.fft_batch_dev(cfg, d_in, d_out, howmany, nfft, nfft)
        torch.cuda.synchronize()
        check(tname, host(d_out), o.fft(x, inverse), nfft, "c2c inv=%d" % inverse)
        assert np.array_equal(host(d_in), x), "input must not be modified"
        # in place
        lib.fft_batch_dev(cfg, d_in, d_in, howmany, nfft, nfft)
        torch.cuda.synchronize()
        check(tname, host(d_in), o.fft(x, inverse), nfft, "c2c in-place inv=%d" % inverse)
        lib.free(cfg)


@pytest.mark.parametrize("nfft", [16, 64, 256, 1000, 1024, 1155, 2048])
def test_generic_kernel_matches_on_fused_sizes(ctx, nfft):
    """the run-time kernel must give the same answer as the compile-time plans (bit-exact in fixed point)"""
    tname, lib, o = ctx
    howmany = 9
    x = random_input(tname, (howmany, nfft), 77 + nfft)
    cfg = lib.alloc(nfft, False)
    d_in, d_out = dev(x), dev(np.zeros_like(x))
    assert lib.plan_kind(nfft) == 1
    lib.force_generic(True)
    try:
        lib.fft_batch_dev(cfg, d_in, d_out, howmany, nfft, nfft)
        torch.cuda.synchronize()
    finally:
        lib.force_generic(False)
    check(tname, host(d_out), o.fft(x, False), nfft, "generic")
    lib.free(cfg)


def test_c2c_strided_input_and_distances(ctx):
    tname, lib, o = ctx
    for nfft, stride in ((64, 3), (1024, 2), (30, 5), (1000, 2)):
        howmany = 6
        x = random_input(tname, (howmany, nfft * stride + 4), 5 + nfft)
        d_in = dev(x)
        d_out = dev(np.zeros((howmany, nfft + 3, 2), x.dtype))
        cfg = lib.alloc(nfft)
        lib.fft_batch_dev(cfg, d_in, d_out, howmany, nfft * stride + 4, nfft + 3, stride)
        torch.cuda.synchronize()
        want = o.fft(x[:, : nfft * stride], False, in_stride=stride, nfft=nfft)
        got = host(d_out)
        check(tname, got[:, :nfft], want, nfft, "stride")
        assert not got[:, nfft:].any(), "padding between output rows must stay untouched"
        lib.free(cfg)


REAL_SIZES = [2, 4, 6, 30, 120, 128, 512, 1000, 2000, 2048, 2310, 4096]


@pytest.mark.parametrize("nfft", REAL_SIZES)
def test_real_batch(ctx, nfft):
    tname, lib, o = ctx
    howmany = 19
    x = random_input(tname, (howmany, nfft), 2000 + nfft, complex_=False)
    nb = nfft // 2 + 1
    d_x, d_X = dev(x), dev(np.zeros((howmany, nb, 2), x.dtype))
    cfg = lib.allocr(nfft, False)
    lib.fftr_batch_dev(cfg, d_x, d_X, howmany, nfft, nb)
    torch.cuda.synchronize()
    want = o.fftr(x)
    check(tname, host(d_X), want, nfft, "fftr")
    lib.free(cfg)
    spec = want if tname in TOL else random_input(tname, (howmany, nb), 2001 + nfft)
    d_S, d_y = dev(spec), dev(np.zeros((howmany, nfft), x.dtype))
    cfgi = lib.allocr(nfft, True)
    lib.fftri_batch_dev(cfgi, d_S, d_y, howmany, nb, nfft)
    torch.cuda.synchronize()
    check(tname, host(d_y), o.fftri(spec), nfft, "fftri")
    lib.free(cfgi)


ND_DIMS = [(8,), (4, 3), (2, 3, 4), (30, 20, 12), (64, 64), (16, 16, 16), (256, 64), (5, 6, 7, 4), (64, 128, 32)]


@pytest.mark.parametrize("inlayout", ["1", "0"])
@pytest.mark.parametrize("dims", ND_DIMS)
def test_fftnd(ctx, dims, inlayout, monkeypatch):
    """both kiss_fftnd strategies: axes transformed where they lie (default) and the reference's transposing sweeps"""
    tname, lib, o = ctx
    monkeypatch.setenv("KISSFFT_FFTND_INLAYOUT", inlayout)
    n = int(np.prod(dims))
    for inverse in (False, True):
        x = random_input(tname, dims, 31 + n)
        want = o.fftnd(x, inverse)
        cfg = lib.allocnd(dims, inverse)
        d_in, d_out = dev(x), dev(np.zeros_like(x))
        lib.fftnd_dev(cfg, d_in, d_out)
        torch.cuda.synchronize()
        check(tname, host(d_out), want, n, "fftnd %s" % (dims,))
        assert np.array_equal(host(d_in), x)
        d_work = dev(np.zeros_like(x))
        lib.fftnd_dev(cfg, d_in, d_in, d_work)     # in place, caller-provided scratch
        torch.cuda.synchronize()
        check(tname, host(d_in), want, n, "fftnd in-place %s" % (dims,))
        lib.free(cfg)


@pytest.mark.parametrize("dims", [(4, 6), (2, 3, 4), (30, 20, 12), (16, 16, 64), (5, 6, 8), (3, 4, 5, 6)])
def test_fftndr(ctx, dims):
    tname, lib, o = ctx
    n = int(np.prod(dims))
    x = random_input(tname, dims, 41 + n, complex_=False)
    want = o.fftndr(x)
    cfg = lib.allocndr(dims, False)
    d_x, d_X = dev(x), dev(np.zeros(want.shape, x.dtype))
    lib.fftndr_dev(cfg, d_x, d_X)
    torch.cuda.synchronize()
    check(tname, host(d_X), want, n, "fftndr %s" % (dims,))
    lib.free(cfg)
    spec = want if tname in TOL else random_input(tname, want.shape[:-1], 43 + n)
    cfgi = lib.allocndr(dims, True)
    d_S, d_y = dev(spec), dev(np.zeros(dims, x.dtype))
    lib.fftndri_dev(cfgi, d_S, d_y)
    torch.cuda.synchronize()
    check(tname, host(d_y), o.fftndri(spec), n, "fftndri %s" % (dims,))
    lib.free(cfgi)


def test_reference_api_with_host_pointers(ctx):
    """the unmodified reference call sequence (kiss_fft_alloc / kiss_fft / free) on ordinary host memory"""
    tname, lib, o = ctx
    nfft = 120
    x = random_input(tname, (nfft,), 3)
    out = np.zeros_like(x)
    cfg = lib.alloc(nfft)
    lib.fft(cfg, x, out)
    check(tname, out, o.fft(x), nfft, "kiss_fft host")
    y = x.copy()
    lib.fft(cfg, y, y)                                    # fin == fout (kiss_fft.c:377-395)
    check(tname, y, o.fft(x), nfft, "kiss_fft host in-place")
    lib.free(cfg)
    xs = random_input(tname, (21 * 5,), 4)
    cfg = lib.alloc(21)
    out = np.zeros((21, 2), xs.dtype)
    lib.fft_stride(cfg, xs, out, 5)
    check(tname, out, o.fft(xs, False, in_stride=5, nfft=21), 21, "kiss_fft_stride host")
    lib.free(cfg)
    # real
    n = 240
    xr = random_input(tname, (n,), 5, complex_=False)
    X = np.zeros((n // 2 + 1, 2), xr.dtype)
    cfg = lib.allocr(n, False)
    lib.fftr(cfg, xr, X)
    check(tname, X, o.fftr(xr), n, "kiss_fftr host")
    before = xr.copy()
    lib.fftri(cfg, X, xr)                                 # wrong direction: logged no-op (kiss_fftr.c:124-127)
    assert np.array_equal(before, xr)
    lib.free(cfg)
    spec = o.fftr(xr) if tname in TOL else random_input(tname, (n // 2 + 1,), 6)
    cfgi = lib.allocr(n, True)
    y = np.zeros(n, xr.dtype)
    lib.fftri(cfgi, spec, y)
    check(tname, y, o.fftri(spec), n, "kiss_fftri host")
    lib.free(cfgi)
    # N-D, also in place as tools/fftutil.c:51 uses it
    dims = (6, 10, 4)
    xn = random_input(tname, dims, 8)
    cfg = lib.allocnd(dims)
    out = np.zeros_like(xn)
    lib.fftnd(cfg, xn, out)
    check(tname, out, o.fftnd(xn), 240, "kiss_fftnd host")
    yn = xn.copy()
    lib.fftnd(cfg, yn, yn)
    check(tname, yn, o.fftnd(xn), 240, "kiss_fftnd host in-place")
    lib.free(cfg)
    dims = (6, 10, 8)
    xr = random_input(tname, dims, 9, complex_=False)
    cfg = lib.allocndr(dims, False)
    want = o.fftndr(xr)
    X = np.zeros(want.shape, xr.dtype)
    lib.fftndr(cfg, xr, X)
    check(tname, X, want, 480, "kiss_fftndr host")
    lib.free(cfg)
    spec = want if tname in TOL else random_input(tname, want.shape[:-1], 10)
    cfgi = lib.allocndr(dims, True)
    y = np.zeros(dims, xr.dtype)
    lib.fftndri(cfgi, spec, y)
    check(tname, y, o.fftndri(spec), 480, "kiss_fftndri host")
    lib.free(cfgi)


@pytest.mark.parametrize("fourstep", ["1", "0"])
@pytest.mark.parametrize("nfft", [16384, 30000, 32768, 65536])
def test_large_lengths(ctx, nfft, fourstep, monkeypatch):
    """lengths beyond the shared-memory kernels: four-step (two fused column passes; float/double with a fused plan for
    both factors) or one radix stage per launch over global memory (fixed point, other lengths, KISSFFT_FOURSTEP=0)"""
    tname, lib, o = ctx
    monkeypatch.setenv("KISSFFT_FOURSTEP", fourstep)
    howmany = 3
    x = random_input(tname, (howmany, nfft), 90 + nfft)
    for inverse in (False, True):
        cfg = lib.alloc(nfft, inverse)
        d_in, d_out = dev(x), dev(np.zeros_like(x))
        lib.fft_batch_dev(cfg, d_in, d_out, howmany, nfft, nfft)
        torch.cuda.synchronize()
        check(tname, host(d_out), o.fft(x, inverse), nfft, "large c2c")
        lib.fft_batch_dev(cfg, d_in, d_in, howmany, nfft, nfft)     # in place
        torch.cuda.synchronize()
        check(tname, host(d_in), o.fft(x, inverse), nfft, "large c2c in place")
        lib.free(cfg)
    n = 2 * nfft
    xr = random_input(tname, (howmany, n), 91 + nfft, complex_=False)
    cfg = lib.allocr(n, False)
    d_x, d_X = dev(xr), dev(np.zeros((howmany, nfft + 1, 2), xr.dtype))
    lib.fftr_batch_dev(cfg, d_x, d_X, howmany, n, nfft + 1)
    torch.cuda.synchronize()
    want = o.fftr(xr)
    check(tname, host(d_X), want, n, "large fftr")
    lib.free(cfg)
    spec = want if tname in TOL else random_input(tname, (howmany, nfft + 1), 92 + nfft)
    cfgi = lib.allocr(n, True)
    d_S, d_y = dev(spec), dev(np.zeros((howmany, n), xr.dtype))
    lib.fftri_batch_dev(cfgi, d_S, d_y, howmany, nfft + 1, n)
    torch.cuda.synchronize()
    check(tname, host(d_y), o.fftri(spec), n, "large fftri")
    lib.free(cfgi)


def test_kfc_cache(ctx):
    """kfc_fft / kfc_ifft (reference kfc.c:63-83, self-test kfc.c:85-108) on host buffers"""
    tname, lib, o = ctx
    for nfft in (512, 360):
        x = random_input(tname, (nfft,), 70 + nfft)
        out = np.zeros_like(x)
        lib.lib.kfc_fft(nfft, x.ctypes.data, out.ctypes.data)
        check(tname, out, o.fft(x), nfft, "kfc_fft")
        back = np.zeros_like(x)
        lib.lib.kfc_ifft(nfft, x.ctypes.data, back.ctypes.data)
        check(tname, back, o.fft(x, True), nfft, "kfc_ifft")
    lib.lib.kfc_cleanup()


def test_host_batch_pipeline(ctx):
    tname, lib, o = ctx
    nfft, howmany = 1024, 300
    x = random_input(tname, (howmany, nfft), 12)
    out = np.zeros_like(x)
    cfg = lib.alloc(nfft)
    lib.fft_batch(cfg, x, out, howmany)
    check(tname, out, o.fft(x), nfft, "kiss_fft_batch host")
    lib.free(cfg)
    n = 4096
    xr = random_input(tname, (howmany, n), 13, complex_=False)
    X = np.zeros((howmany, n // 2 + 1, 2), xr.dtype)
    cfg = lib.allocr(n, False)
    lib.fftr_batch(cfg, xr, X, howmany)
    want = o.fftr(xr)
    check(tname, X, want, n, "kiss_fftr_batch host")
    lib.free(cfg)
    spec = want if tname in TOL else random_input(tname, (howmany, n // 2 + 1), 14)
    y = np.zeros((howmany, n), xr.dtype)
    cfgi = lib.allocr(n, True)
    lib.fftri_batch(cfgi, spec, y, howmany)
    check(tname, y, o.fftri(spec), n, "kiss_fftri_batch host")
    lib.free(cfgi)


def test_structured_inputs(ctx):
    """impulse, constant and two-tone rows (reference test/twotonetest.c:38-47)"""
    tname, lib, o = ctx
    nfft = 1024
    amp = 1.0 if tname in TOL else (16383 if tname == "int16_t" else 1073741823)
    rows = np.zeros((4, nfft, 2), np.float64)
    rows[0, 0, 0] = amp
    rows[1, :, 0] = amp
    t = np.arange(nfft)
    rows[2, :, 0] = amp * 0.5 * (np.cos(2 * np.pi * 17 * t / nfft) + np.cos(2 * np.pi * 200 * t / nfft))
    rows[3, 5, 1] = -amp
    x = rows.astype(o.dtype)
    cfg = lib.alloc(nfft)
    d_in, d_out = dev(x), dev(np.zeros_like(x))
    lib.fft_batch_dev(cfg, d_in, d_out, 4, nfft, nfft)
    torch.cuda.synchronize()
    check(tname, host(d_out), o.fft(x), nfft, "structured")
    lib.free(cfg)


def test_error_paths(ctx):
    tname, lib, o = ctx
    import kissfft_b200
    assert lib.lib.kiss_fftr_alloc(31, 0, None, None) is None          # odd real length (kiss_fftr.c:29-32)
    need = ctypes.c_size_t(0)
    assert lib.lib.kiss_fft_alloc(64, 0, None, ctypes.byref(need)) is None and need.value > 0   # size query
    small = ctypes.c_size_t(8)
    buf = ctypes.create_string_buffer(need.value)
    assert lib.lib.kiss_fft_alloc(64, 0, buf, ctypes.byref(small)) is None and small.value == need.value
    ok = ctypes.c_size_t(need.value)
    cfg = lib.lib.kiss_fft_alloc(64, 0, buf, ctypes.byref(ok))
    assert cfg == ctypes.addressof(buf)
    x = random_input(tname, (2, 64), 1)
    d_in, d_out = dev(x), dev(np.zeros_like(x))
    lib.fft_batch_dev(cfg, d_in, d_out, 2, 64, 64)                      # a cfg placed in caller memory works
    torch.cuda.synchronize()
    check(tname, host(d_out), o.fft(x), 64, "placed cfg")
    cfgr = lib.allocr(64, True)
    with pytest.raises(kissfft_b200.KissFFTError):
        lib.fftr_batch_dev(cfgr, d_in, d_out, 1, 64, 33)               # forward call with an inverse cfg
    lib.free(cfgr)


# ---- BASELINE.json configurations at full size ----------------------------------------------------------------

def _full_batch_check(tname, lib, o, nfft, howmany, seed, sample=1024):
    """full batch on the GPU; oracle on an evenly spaced sample of rows + whole-batch size-independent properties"""
    x = random_input(tname, (howmany, nfft), seed)
    d_in = dev(x)
    d_out = torch.zeros_like(d_in)
    cfg = lib.alloc(nfft)
    lib.fft_batch_dev(cfg, d_in, d_out, howmany, nfft, nfft)
    torch.cuda.synchronize()
    got = host(d_out)
    idx = np.unique(np.concatenate([np.linspace(0, howmany - 1, sample).astype(np.int64), np.arange(howmany - 8, howmany)]))
    check(tname, got[idx], o.fft(x[idx]), nfft, "full-size sample")
    if have_reference(tname):
        # the compiled reference itself over the WHOLE batch, all host cores
        ref = Reference(tname).fft(x, nthreads=0 or __import__("os").cpu_count())
        check(tname, got, ref, nfft, "full batch vs compiled reference")
    if tname in TOL:
        # Parseval over the whole batch: sum |X|^2 == N * sum |x|^2
        ex = np.sum(x.astype(np.float64) ** 2, axis=(1, 2))
        eX = np.sum(got.astype(np.float64) ** 2, axis=(1, 2))
        assert np.allclose(eX, nfft * ex, rtol=1e-4 if tname == "float" else 1e-10)
        # forward o inverse == N * x  (reference README: unscaled in both directions)
        cfgi = lib.alloc(nfft, True)
        d_back = torch.zeros_like(d_in)
        lib.fft_batch_dev(cfgi, d_out, d_back, howmany, nfft, nfft)
        torch.cuda.synchronize()
        assert rel_rms(host(d_back) / nfft, x) <= 2 * TOL[tname] * np.log2(nfft)
        lib.free(cfgi)
    else:
        # DC bin == sround-chain of the row sum: cheap whole-batch property -- every row's output must be
        # identical when the same row is transformed alone (batch independence)
        pick = np.array([0, howmany // 3, howmany - 1])
        d_one = dev(x[pick])
        d_res = torch.zeros_like(d_one)
        lib.fft_batch_dev(cfg, d_one, d_res, len(pick), nfft, nfft)
        torch.cuda.synchronize()
        assert np.array_equal(host(d_res), got[pick])
    lib.free(cfg)


def test_config1_c2c_1024_x_65536():
    import kissfft_b200
    _full_batch_check("float", kissfft_b200.get("float"), Oracle("float"), 1024, 65536, 11)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft", [1000, 1155])
def test_config3_mixed_radix_x_100000(tname, nfft):
    import kissfft_b200
    _full_batch_check(tname, kissfft_b200.get(tname), Oracle(tname), nfft, 100000, 12, sample=512)


@pytest.mark.parametrize("tname", ["int16_t", "int32_t"])
def test_config4_fixed_2048_x_65536(tname):
    import kissfft_b200
    _full_batch_check(tname, kissfft_b200.get(tname), Oracle(tname), 2048, 65536, 13, sample=512)


def test_config2_real_4096_x_32768():
    import kissfft_b200
    tname, nfft, howmany = "float", 4096, 32768
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    x = random_input(tname, (howmany, nfft), 14, complex_=False)
    nb = nfft // 2 + 1
    d_x = dev(x)
    d_X = torch.zeros((howmany, nb, 2), dtype=d_x.dtype, device="cuda")
    cfg = lib.allocr(nfft, False)
    lib.fftr_batch_dev(cfg, d_x, d_X, howmany, nfft, nb)
    torch.cuda.synchronize()
    got = host(d_X)
    idx = np.linspace(0, howmany - 1, 512).astype(np.int64)
    check(tname, got[idx], o.fftr(x[idx]), nfft, "r2c sample")
    if have_reference(tname):
        check(tname, got, Reference(tname).fftr(x, nthreads=__import__("os").cpu_count()), nfft, "r2c full vs reference")
    # C2R(R2C(x)) == nfft * x over the whole batch
    cfgi = lib.allocr(nfft, True)
    d_y = torch.zeros_like(d_x)
    lib.fftri_batch_dev(cfgi, d_X, d_y, howmany, nb, nfft)
    torch.cuda.synchronize()
    assert rel_rms(host(d_y) / nfft, x) <= 2e-6 * np.log2(nfft)
    check(tname, host(d_y)[idx], o.fftri(got[idx]), nfft, "c2r sample")
    if have_reference(tname):
        # kiss_fftri of the compiled reference on the same spectra, the WHOLE batch (like R2C above)
        check(tname, host(d_y), Reference(tname).fftri(got, nthreads=__import__("os").cpu_count()), nfft, "c2r full vs reference")
    lib.free(cfg)
    lib.free(cfgi)


def test_config5_3d_single_gpu_256():
    """kiss_fftnd 3-D at 256^3 against the oracle (1024^3 is exercised by bench.py / the multi-GPU tests)"""
    import kissfft_b200
    tname = "float"
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    dims = (256, 256, 256)
    x = random_input(tname, dims, 15)
    d_in = dev(x)
    d_out = torch.zeros_like(d_in)
    d_work = torch.zeros_like(d_in)
    cfg = lib.allocnd(dims)
    lib.fftnd_dev(cfg, d_in, d_out, d_work)
    torch.cuda.synchronize()
    got = host(d_out)
    want = np.fft.fftn(x[..., 0].astype(np.float64) + 1j * x[..., 1].astype(np.float64))
    assert rel_rms(got, np.stack([want.real, want.imag], -1)) <= 1e-6 * 24
    if have_reference(tname):
        check(tname, got, Reference(tname).fftnd(x), 256 ** 3, "3-D vs compiled reference")
    else:
        check(tname, got, o.fftnd(x), 256 ** 3, "3-D vs oracle")
    lib.free(cfg)


def sampled_dft_check(x, X, nbins, seed, tol):
    """SURVEY.md 8(d) parity sampling for arrays too large for the CPU oracle: Parseval over the whole array and `nbins`
    random output bins against direct DFT sums accumulated in float64 (on the GPU, chunked over the leading axis).
    x, X: torch complex-as-(...,2) float32 tensors of shape dims + (2,), X = DFT(x) in natural order."""
    dims = tuple(x.shape[:-1])
    d0, rest = dims[0], int(np.prod(dims[1:]))
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    ks = [torch.randint(0, d, (nbins,), generator=g) for d in dims]
    cd = torch.complex128
    # separable phase factors exp(-2 pi i k_a n_a / d_a), one column per sampled bin
    W = [torch.exp(-2j * np.pi * (torch.arange(d, dtype=torch.float64)[:, None] * k[None, :].double()) / d).to(cd).cuda()
         for d, k in zip(dims, ks)]
    acc = torch.zeros(nbins, dtype=cd, device="cuda")
    ein = torch.zeros((), dtype=torch.float64, device="cuda")
    eout = torch.zeros((), dtype=torch.float64, device="cuda")
    step = max(1, (1 << 25) // rest)
    for p0 in range(0, d0, step):
        p1 = min(d0, p0 + step)
        xc = torch.view_as_complex(x[p0:p1].contiguous()).to(cd)             # [p][d1]...[dlast]
        ein += (xc.real ** 2 + xc.imag ** 2).sum()
        Xc = X[p0:p1].double()
        eout += (Xc ** 2).sum()
        del Xc
        t = xc.reshape(-1, dims[-1]) @ W[-1]                                  # contract the last axis -> [.., nbins]
        del xc
        for a in range(len(dims) - 2, 0, -1):                                 # then the middle axes, innermost first
            t = (t.reshape(-1, dims[a], nbins) * W[a][None]).sum(1)
        acc += (t.reshape(p1 - p0, nbins) * W[0][p0:p1]).sum(0)
    n = float(np.prod(dims))
    assert abs(float(eout) / (n * float(ein)) - 1.0) <= tol, "Parseval: %r vs %r" % (float(eout), n * float(ein))
    idx = tuple(k.cuda() for k in ks)
    got = torch.view_as_complex(X[idx].contiguous()).to(cd)
    err = float((got - acc).abs().pow(2).sum().sqrt() / acc.abs().pow(2).sum().sqrt())
    assert err <= tol, "sampled bins rel-rms %.3g > %.3g" % (err, tol)
    return err


def test_config5_3d_single_gpu_1024():
    """BASELINE configs[4] on one GPU: kiss_fftnd 1024^3 complex float (8 GiB in, 8 GiB out, no work buffer).  The CPU
    oracle would need ~5 min and 24 GiB here, so SURVEY.md 8(d)'s sampling applies: Parseval over the whole array +
    64 random bins vs float64 direct DFT sums, tolerance 1e-6*log2(N)."""
    import kissfft_b200
    free, _ = torch.cuda.mem_get_info()
    if free < 19 * (1 << 30):
        pytest.skip("needs 19 GiB of free device memory")
    lib = kissfft_b200.get("float")
    dims = (1024, 1024, 1024)
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.rand(dims + (2,), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    X = torch.empty_like(x)
    cfg = lib.allocnd(dims)
    lib.fftnd_dev(cfg, x, X, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    sampled_dft_check(x, X, 64, 77, 1e-6 * 30)
    # inverse of the forward result returns N * x (kiss_fftnd does not scale): whole-array round trip, in place
    cfgi = lib.allocnd(dims, True)
    lib.fftnd_dev(cfgi, X, X, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    num = torch.zeros((), dtype=torch.float64, device="cuda")
    den = torch.zeros((), dtype=torch.float64, device="cuda")
    for p0 in range(0, 1024, 64):
        d = X[p0:p0 + 64].double() / float(1 << 30) - x[p0:p0 + 64].double()
        num += (d ** 2).sum()
        den += (x[p0:p0 + 64].double() ** 2).sum()
    assert float((num / den).sqrt()) <= 2e-6 * 30
    lib.free(cfg)
    lib.free(cfgi)


def test_slab_single_rank_and_planes_pass():
    """slab decomposition with G == 1 (steps A, B, C without the exchange) and the plane-batched column pass"""
    import kissfft_b200
    from kissfft_b200.slab import SlabFFT3D
    tname = "float"
    lib = kissfft_b200.get(tname)
    dims = (64, 32, 128)
    x = random_input(tname, dims, 21)
    plan = SlabFFT3D(dims, tname=tname)
    d_x, send, recv, out = plan.alloc()
    d_x.copy_(torch.from_numpy(x))
    plan.forward(d_x, send, recv, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    nat = host(plan.gather_natural(out))
    want = np.fft.fftn(x[..., 0].astype(np.float64) + 1j * x[..., 1].astype(np.float64))
    assert rel_rms(nat, np.stack([want.real, want.imag], -1)) <= 1e-6 * np.log2(x.size / 2)
    # planes pass alone, non-power-of-two plane shape -> run-time kernel
    P, d1, d2 = 3, 30, 20
    y = random_input(tname, (P, d1, d2), 22)
    cfg = lib.alloc(d1)
    d_y, d_o = dev(y), dev(np.zeros((P, d2, d1, 2), y.dtype))
    lib.planes_pass_dev(cfg, d_y, d_o, P, d2, d2, d1 * d2, d2 * d1)
    torch.cuda.synchronize()
    o = Oracle(tname)
    want = np.stack([o.fft(np.ascontiguousarray(y[p].transpose(1, 0, 2))) for p in range(P)])
    check(tname, host(d_o), want, d1, "planes pass")
    lib.free(cfg)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft,nimp", [(256, 33), (512, 100), (1024, 129), (2048, 513), (4096, 1000), (0, 200)])
def test_fused_fast_convolution(tname, nfft, nimp):
    """overlap-scrap FIR filtering (reference tools/kiss_fastfir.c) fused into one kernel vs the same pipeline assembled
    from oracle transforms: H = FFT(rotated impulse response)/nfft, per block IFFT(FFT(x) .* H), first ngood samples kept"""
    import kissfft_b200
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    rng = np.random.default_rng(3)
    imp = rng.uniform(-1, 1, size=(nimp, 2)).astype(o.dtype) / nimp
    cfg, n, ngood = lib.fastconv_alloc(imp, nfft)
    assert ngood == n - nimp + 1 and (nfft == 0 or n == nfft)
    nblocks = 37
    nsamp = (nblocks - 1) * ngood + n + 5                      # 5 trailing samples that do not complete a block
    x = random_input(tname, (nsamp,), 17)
    d_in, d_out = dev(x), dev(np.zeros_like(x))
    done = lib.fastconv_dev(cfg, d_in, d_out, nsamp)
    torch.cuda.synchronize()
    assert done == nblocks * ngood
    # oracle pipeline
    rot = np.zeros((n, 2), o.dtype)
    rot[0] = imp[nimp - 1]
    rot[n - nimp + 1:] = imp[: nimp - 1]
    H = o.fft(rot) * np.float32(1.0 / n)
    blocks = np.stack([x[b * ngood: b * ngood + n] for b in range(nblocks)])
    X = o.fft(blocks).astype(np.float64)
    Hc = H.astype(np.float64)
    Y = np.empty_like(X)
    Y[..., 0] = X[..., 0] * Hc[:, 0] - X[..., 1] * Hc[:, 1]
    Y[..., 1] = X[..., 0] * Hc[:, 1] + X[..., 1] * Hc[:, 0]
    y = o.fft(Y.astype(o.dtype), True)[:, :ngood].reshape(-1, 2)
    got = host(d_out)
    assert rel_rms(got[: done], y) <= 3 * TOL[tname] * np.log2(n)
    assert not got[done:].any(), "samples beyond the processed range must stay untouched"
    # and against the textbook answer: linear convolution with the impulse response
    xc, hc = x[:, 0].astype(np.float64) + 1j * x[:, 1], imp[:, 0].astype(np.float64) + 1j * imp[:, 1]
    full = np.convolve(xc, hc)[nimp - 1: nimp - 1 + done]
    assert rel_rms(got[:done], np.stack([full.real, full.imag], -1)) <= 20 * TOL[tname] * np.log2(n)
    lib.fastconv_free(cfg)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft,nimp", [(1000, 101), (300, 7), (8192, 2500), (0, 3000), (32768, 4097)])
def test_fast_convolution_any_length(tname, nfft, nimp):
    """lengths without a fused kernel (the reference's kiss_fastfir works for any size; its default rule gives 8192 for
    filters over 2048 taps): composed from block gather + batched transforms + pointwise product, vs np.convolve"""
    import kissfft_b200
    lib = kissfft_b200.get(tname)
    rng = np.random.default_rng(nimp)
    imp = (rng.random((nimp, 2)) - 0.5).astype(lib.dtype)
    cfg, n, ngood = lib.fastconv_alloc(imp, nfft)
    assert n >= nimp and ngood == n - nimp + 1
    nsamp = 5 * ngood + n + 3
    x = (rng.random((nsamp, 2)) * 2 - 1).astype(lib.dtype)
    d_in = dev(x)
    d_out = torch.zeros_like(d_in)
    done = lib.fastconv_dev(cfg, d_in, d_out, nsamp, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert done == ((nsamp - n) // ngood + 1) * ngood
    got = host(d_out)
    assert not got[done:].any()
    xc, hc = x[:, 0].astype(np.float64) + 1j * x[:, 1], imp[:, 0].astype(np.float64) + 1j * imp[:, 1]
    # the reference scales the response by a FLOAT 1/nfft in every build (kiss_fastfir.c:72,152): exact for powers of two only
    full = np.convolve(xc, hc)[nimp - 1: nimp - 1 + done] * (float(np.float32(1.0 / n)) * n)
    assert rel_rms(got[:done], np.stack([full.real, full.imag], -1)) <= 20 * TOL[tname] * np.log2(n)
    lib.fastconv_free(cfg)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft,nimp", [(256, 33), (4096, 1000), (1000, 100), (0, 200), (16384, 5000)])
def test_real_fast_convolution(tname, nfft, nimp):
    """the reference's REAL_FASTFIR build (tools/kiss_fastfir.c:26-33, 91-95): real samples through kiss_fftr / kiss_fftri;
    odd block advances (ngood) included.  Checked against the oracle's own fftr -> multiply -> fftri pipeline on one block
    and against np.convolve on all of them."""
    import kissfft_b200
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    rng = np.random.default_rng(nimp + 1)
    imp = (rng.random(nimp) - 0.5).astype(lib.dtype)
    cfg, n, ngood = lib.fastconvr_alloc(imp, nfft)
    nsamp = 4 * ngood + n + 5
    x = (rng.random(nsamp) * 2 - 1).astype(lib.dtype)
    d_in = dev(x)
    d_out = torch.zeros_like(d_in)
    done = lib.fastconvr_dev(cfg, d_in, d_out, nsamp, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert done == ((nsamp - n) // ngood + 1) * ngood
    got = host(d_out)
    assert not got[done:].any()
    full = np.convolve(x.astype(np.float64), imp.astype(np.float64))[nimp - 1: nimp - 1 + done] * (float(np.float32(1.0 / n)) * n)
    assert rel_rms(got[:done], full) <= 20 * TOL[tname] * np.log2(n)
    # first block against the reference's arithmetic: fftr(block) * (fftr(rotated response) / nfft) -> fftri
    rot = np.zeros(n, lib.dtype)
    rot[0] = imp[nimp - 1]
    rot[n - nimp + 1:] = imp[:nimp - 1]
    H = o.fftr(rot[None])[0] * lib.dtype(np.float32(1.0 / n))
    X = o.fftr(x[None, :n])[0]
    Y = np.stack([X[:, 0] * H[:, 0] - X[:, 1] * H[:, 1], X[:, 0] * H[:, 1] + X[:, 1] * H[:, 0]], -1).astype(lib.dtype)
    want = o.fftri(Y[None])[0][:ngood]
    assert rel_rms(got[:ngood], want) <= 4 * TOL[tname] * np.log2(n)
    lib.fastconv_free(cfg)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("nfft", [16384, 65536, 1 << 20])
def test_fourstep_long_rows(tname, nfft, monkeypatch):
    """default for float/double (KISSFFT_FOURSTEP=0 opts out): long contiguous rows as two fused column passes instead of one launch per radix stage"""
    import torch
    import kissfft_b200
    from oracle.loader import Oracle, random_input, rel_rms
    monkeypatch.delenv("KISSFFT_FOURSTEP", raising=False)
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    rows = 3
    x = random_input(tname, (rows, nfft), 4242)
    for inverse in (False, True):
        cfg = lib.alloc(nfft, inverse)
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.empty_like(d_in)
        before = lib.launch_count()
        lib.fft_batch_dev(cfg, d_in, d_out, rows, nfft, nfft, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert lib.launch_count() - before == 2
        tol = (1e-6 if tname == "float" else 1e-14) * np.log2(nfft)
        assert rel_rms(d_out.cpu().numpy(), o.fft(x, inverse)) <= tol


@pytest.mark.parametrize("tname", ["float", "double", "int16_t", "int32_t"])
@pytest.mark.parametrize("dims", [(64, 128, 256), (128, 64), (256, 256, 30)])
def test_fftnd_in_layout(tname, dims, monkeypatch):
    """default (KISSFFT_FFTND_INLAYOUT=0 opts out): every axis transformed where it lies; same bits as the transposing sweeps in fixed point"""
    import torch
    import kissfft_b200
    from oracle.loader import Oracle, random_input, rel_rms
    monkeypatch.delenv("KISSFFT_FFTND_INLAYOUT", raising=False)
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    x = random_input(tname, dims, 99)
    cfg = lib.allocnd(list(dims), False)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty_like(d_in)
    lib.fftnd_dev(cfg, d_in, d_out, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = o.fftnd(x)
    if tname in ("float", "double"):
        n = float(np.prod(dims))
        assert rel_rms(d_out.cpu().numpy(), want) <= (1e-6 if tname == "float" else 1e-14) * np.log2(n)
    else:
        assert np.array_equal(d_out.cpu().numpy(), want)
    # in place
    lib.fftnd_dev(cfg, d_in, d_in, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(d_in, d_out)


@pytest.mark.parametrize("tname", ["float", "double"])
@pytest.mark.parametrize("work", ["caller", "internal", "host"])
@pytest.mark.parametrize("dims", [(2, 3, 4), (30, 20, 12), (64, 128, 32), (16, 1024, 48), (1024, 16, 32), (8, 40, 1024), (256, 256, 64)])
def test_fftnd_permuted_passes(tname, dims, work, monkeypatch):
    """3-D float/double: the three plane-local transposing passes (axis 1, 0, 2 with permuted row placement,
    kf_api.c:kf_fftnd3_permuted) against the oracle's kiss_fftnd -- forced on for small arrays, forward and inverse,
    with a caller-provided work buffer, the library's own, and through the host-pointer kiss_fftnd"""
    import torch
    import kissfft_b200
    from oracle.loader import Oracle, random_input, rel_rms
    monkeypatch.setenv("KISSFFT_FFTND_PERMUTE", "1")
    lib, o = kissfft_b200.get(tname), Oracle(tname)
    n = float(np.prod(dims))
    tol = (1e-6 if tname == "float" else 1e-14) * max(np.log2(n), 1.0)
    for inverse in (False, True):
        x = random_input(tname, dims, 7 + int(inverse))
        want = o.fftnd(x, inverse)
        cfg = lib.allocnd(list(dims), inverse)
        if work == "host":
            got = np.zeros_like(x)
            lib.fftnd(cfg, x, got)
        else:
            d_in = torch.from_numpy(x).cuda()
            d_out = torch.zeros_like(d_in)
            d_work = torch.zeros_like(d_in) if work == "caller" else None
            n0 = lib.launch_count()
            lib.fftnd_dev(cfg, d_in, d_out, d_work, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert 3 <= lib.launch_count() - n0 <= 6, "three passes expected (a ragged tail of a pass may add a launch)"
            assert torch.equal(d_in.cpu(), torch.from_numpy(x)), "the input is const"
            got = d_out.cpu().numpy()
        assert rel_rms(got, want) <= tol
        lib.free(cfg)
