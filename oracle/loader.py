"""ctypes front-end to the parity checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this module;
the product package kissfft_b200 never does.

Two checkers, same numpy-level interface (`Oracle`, `Reference`):
  * Oracle    -- oracle/_lib/liboracle-<type>.so, our C restatement (kiss_oracle.c); built on demand with gcc.
  * Reference -- oracle/_ref/libkissfft-<type>.so, the unmodified reference compiled by `make -C oracle ref`
                 (only buildable where /root/reference exists; the built .so travels to the GPU box).

Complex data is always an ndarray of the scalar dtype with a trailing axis of 2 (r, i) so that the four
datatypes (float32, float64, int16 = Q15, int32 = Q31) are handled uniformly.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TYPES = ("float", "double", "int16_t", "int32_t")
NP_DTYPE = {"float": np.float32, "double": np.float64, "int16_t": np.int16, "int32_t": np.int32}
SAMP_MAX = {"int16_t": 32767, "int32_t": 2147483647}
REF_SRC = os.environ.get("KISSFFT_REFERENCE", "/root/reference")

K_FFT, K_FFTR, K_FFTRI, K_FFTND, K_FFTNDR, K_FFTNDRI = range(6)


def _make(*targets):
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run(["make", "-s", "-C", HERE, "GCC=" + gcc, "REF=" + REF_SRC, *targets], check=True)


def build_oracle():
    """Compile the C restatement + CPU batch driver (always possible: needs only gcc)."""
    _make("all")


def build_reference():
    """Compile the unmodified reference into oracle/_ref (needs /root/reference). Returns True if built."""
    if not os.path.isdir(REF_SRC):
        return False
    _make("ref")
    return True


def build_dropin():
    """Compile the reference's own unmodified test/tool programs against THIS repo's headers and libraries
    (oracle/_ref/dropin/*; needs /root/reference and the built kissfft_b200/lib).  Returns True if built."""
    if not os.path.isdir(REF_SRC):
        return False
    _make("dropin")
    return True


def dropin_path(name, tname):
    return os.path.join(HERE, "_ref", "dropin", "%s-%s" % (name, tname))


def oracle_lib_path(tname):
    return os.path.join(HERE, "_lib", "liboracle-%s.so" % tname)


def reference_lib_path(tname, openmp=False):
    return os.path.join(HERE, "_ref", "libkissfft-%s%s.so" % (tname, "-openmp" if openmp else ""))


def have_reference(tname="float"):
    return os.path.exists(reference_lib_path(tname))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Oracle:
    """numpy interface to liboracle-<type>.so (functions of kiss_oracle.c)."""

    def __init__(self, tname):
        assert tname in TYPES
        path = oracle_lib_path(tname)
        if not os.path.exists(path):
            build_oracle()
        self.tname = tname
        self.dtype = NP_DTYPE[tname]
        self.lib = ctypes.CDLL(path)
        L = self.lib
        vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.oracle_fft_batch.argtypes = [ci, ci, vp, vp, ci, sz, sz, sz]
        L.oracle_fft_batch.restype = None
        L.oracle_fftr_batch.argtypes = [ci, vp, vp, sz]
        L.oracle_fftri_batch.argtypes = [ci, vp, vp, sz]
        L.oracle_fftnd.argtypes = [vp, ci, ci, vp, vp]
        L.oracle_fftnd.restype = None
        L.oracle_fftndr.argtypes = [vp, ci, vp, vp]
        L.oracle_fftndri.argtypes = [vp, ci, vp, vp]
        L.oracle_factor.argtypes = [ci, vp]
        L.oracle_twiddles.argtypes = [ci, ci, vp]
        L.oracle_twiddles.restype = None
        L.oracle_super_twiddles.argtypes = [ci, ci, vp]
        L.oracle_super_twiddles.restype = None
        assert L.oracle_sizeof_scalar() == np.dtype(self.dtype).itemsize

    def factor(self, n):
        buf = np.zeros(64, np.int32)
        ns = self.lib.oracle_factor(n, _ptr(buf))
        return [(int(buf[2 * s]), int(buf[2 * s + 1])) for s in range(ns)]

    def twiddles(self, nfft, inverse):
        tw = np.empty((nfft, 2), self.dtype)
        self.lib.oracle_twiddles(nfft, int(inverse), _ptr(tw))
        return tw

    def super_twiddles(self, ncfft, inverse):
        st = np.empty((ncfft // 2, 2), self.dtype)
        self.lib.oracle_super_twiddles(ncfft, int(inverse), _ptr(st))
        return st

    def fft(self, x, inverse=False, in_stride=1, nfft=None):
        """x: (howmany, nfft*in_stride, 2) or (nfft*in_stride, 2) -> same leading shape, (.., nfft, 2)."""
        x = _c(x, self.dtype)
        single = x.ndim == 2
        xb = x[None] if single else x
        howmany, span = xb.shape[0], xb.shape[1]
        if nfft is None:
            nfft = span // in_stride
        out = np.empty((howmany, nfft, 2), self.dtype)
        self.lib.oracle_fft_batch(nfft, int(inverse), _ptr(xb), _ptr(out), in_stride, howmany, span, nfft)
        return out[0] if single else out

    def fftr(self, x):
        """x: (howmany, nfft) real -> (howmany, nfft/2+1, 2)."""
        x = _c(x, self.dtype)
        single = x.ndim == 1
        xb = x[None] if single else x
        howmany, nfft = xb.shape
        out = np.empty((howmany, nfft // 2 + 1, 2), self.dtype)
        rc = self.lib.oracle_fftr_batch(nfft, _ptr(xb), _ptr(out), howmany)
        if rc != 0:
            raise ValueError("real FFT length must be even")
        return out[0] if single else out

    def fftri(self, X):
        """X: (howmany, nfft/2+1, 2) -> (howmany, nfft) real."""
        X = _c(X, self.dtype)
        single = X.ndim == 2
        Xb = X[None] if single else X
        howmany, nb = Xb.shape[0], Xb.shape[1]
        nfft = 2 * (nb - 1)
        out = np.empty((howmany, nfft), self.dtype)
        rc = self.lib.oracle_fftri_batch(nfft, _ptr(Xb), _ptr(out), howmany)
        if rc != 0:
            raise ValueError("real FFT length must be even")
        return out[0] if single else out

    def fftnd(self, x, inverse=False):
        """x: (d0, d1, ..., 2)."""
        x = _c(x, self.dtype)
        dims = np.array(x.shape[:-1], np.int32)
        out = np.empty_like(x)
        self.lib.oracle_fftnd(_ptr(dims), len(dims), int(inverse), _ptr(x), _ptr(out))
        return out

    def fftndr(self, x):
        x = _c(x, self.dtype)
        dims = np.array(x.shape, np.int32)
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 2 + 1, 2), self.dtype)
        rc = self.lib.oracle_fftndr(_ptr(dims), len(dims), _ptr(x), _ptr(out))
        if rc != 0:
            raise ValueError("last dimension must be even")
        return out

    def fftndri(self, X):
        X = _c(X, self.dtype)
        dims = np.array(X.shape[:-2] + (2 * (X.shape[-2] - 1),), np.int32)
        out = np.empty(tuple(dims), self.dtype)
        rc = self.lib.oracle_fftndri(_ptr(dims), len(dims), _ptr(X), _ptr(out))
        if rc != 0:
            raise ValueError("last dimension must be even")
        return out


class CpuDriver:
    """oracle/cpu_driver.c: loops any kissfft-API library over a batch, optionally with OpenMP threads."""

    def __init__(self):
        path = os.path.join(HERE, "_lib", "libcpudrv.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = ctypes.CDLL(path)
        vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        self.lib.cpudrv_run.argtypes = [ctypes.c_char_p, ci, vp, ci, ci, vp, vp, sz, sz, sz, ci, ci]
        self.lib.cpudrv_run.restype = ctypes.c_double
        self.lib.cpudrv_max_threads.restype = ci

    def max_threads(self):
        return int(self.lib.cpudrv_max_threads())

    def run(self, libpath, kind, dims, inverse, inp, out, howmany, in_dist_bytes, out_dist_bytes, nthreads=1, reps=1):
        dims = np.array(dims, np.int32)
        t = self.lib.cpudrv_run(libpath.encode(), kind, _ptr(dims), len(dims), int(inverse), _ptr(inp), _ptr(out),
                                howmany, in_dist_bytes, out_dist_bytes, nthreads, reps)
        if t < 0:
            raise RuntimeError("cpudrv_run failed (%g) for %s" % (t, libpath))
        return t


class Reference:
    """numpy interface to the compiled UNMODIFIED reference (oracle/_ref/libkissfft-<type>.so)."""

    def __init__(self, tname, openmp=False):
        assert tname in TYPES
        self.path = reference_lib_path(tname, openmp)
        if not os.path.exists(self.path):
            if not build_reference():
                raise FileNotFoundError(self.path)
        self.tname = tname
        self.dtype = NP_DTYPE[tname]
        self.esz = np.dtype(self.dtype).itemsize
        self.drv = CpuDriver()

    def _run(self, kind, dims, inverse, inp, out, howmany, ind, outd, nthreads=1):
        self.drv.run(self.path, kind, dims, inverse, inp, out, howmany, ind, outd, nthreads)

    def fft(self, x, inverse=False, nthreads=1):
        x = _c(x, self.dtype)
        single = x.ndim == 2
        xb = x[None] if single else x
        howmany, nfft = xb.shape[0], xb.shape[1]
        out = np.empty_like(xb)
        self._run(K_FFT, [nfft], inverse, xb, out, howmany, nfft * 2 * self.esz, nfft * 2 * self.esz, nthreads)
        return out[0] if single else out

    def fft_stride(self, x, nfft, in_stride, inverse=False):
        """single transform through kiss_fft_stride (direct ctypes call)."""
        x = _c(x, self.dtype)
        lib = ctypes.CDLL(self.path, mode=ctypes.RTLD_LOCAL)
        lib.kiss_fft_alloc.restype = ctypes.c_void_p
        lib.kiss_fft_alloc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.kiss_fft_stride.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.kiss_fft_stride.restype = None
        cfg = lib.kiss_fft_alloc(nfft, int(inverse), None, None)
        out = np.empty((nfft, 2), self.dtype)
        lib.kiss_fft_stride(cfg, _ptr(x), _ptr(out), in_stride)
        ctypes.CDLL(None).free(ctypes.c_void_p(cfg))
        return out

    def fftr(self, x, nthreads=1):
        x = _c(x, self.dtype)
        single = x.ndim == 1
        xb = x[None] if single else x
        howmany, nfft = xb.shape
        out = np.empty((howmany, nfft // 2 + 1, 2), self.dtype)
        self._run(K_FFTR, [nfft], 0, xb, out, howmany, nfft * self.esz, (nfft // 2 + 1) * 2 * self.esz, nthreads)
        return out[0] if single else out

    def fftri(self, X, nthreads=1):
        X = _c(X, self.dtype)
        single = X.ndim == 2
        Xb = X[None] if single else X
        howmany, nb = Xb.shape[0], Xb.shape[1]
        nfft = 2 * (nb - 1)
        out = np.empty((howmany, nfft), self.dtype)
        self._run(K_FFTRI, [nfft], 1, Xb, out, howmany, nb * 2 * self.esz, nfft * self.esz, nthreads)
        return out[0] if single else out

    def fftnd(self, x, inverse=False):
        x = _c(x, self.dtype)
        out = np.empty_like(x)
        self._run(K_FFTND, list(x.shape[:-1]), inverse, x, out, 1, 0, 0)
        return out

    def fftndr(self, x):
        x = _c(x, self.dtype)
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 2 + 1, 2), self.dtype)
        self._run(K_FFTNDR, list(x.shape), 0, x, out, 1, 0, 0)
        return out

    def fftndri(self, X):
        X = _c(X, self.dtype)
        dims = list(X.shape[:-2]) + [2 * (X.shape[-2] - 1)]
        out = np.empty(tuple(dims), self.dtype)
        self._run(K_FFTNDRI, dims, 1, X, out, 1, 0, 0)
        return out


# ---- seeded inputs shared by tests, smoke and bench --------------------------------------------------------

def random_input(tname, shape, seed, complex_=True):
    """i.i.d. uniform(-1,1) for float/double (reference test/testkiss.py:54-59); uniform integers in
    [-SAMP_MAX/2, SAMP_MAX/2] for Q15/Q31 (the range of reference test/test_real.c:21-30, where the
    fixed-point outputs are compiler-independent -- SURVEY.md section 8c)."""
    rng = np.random.default_rng(seed)
    full = tuple(shape) + ((2,) if complex_ else ())
    if tname in ("float", "double"):
        return rng.uniform(-1.0, 1.0, size=full).astype(NP_DTYPE[tname])
    half = SAMP_MAX[tname] // 2
    return rng.integers(-half, half + 1, size=full, dtype=np.int64).astype(NP_DTYPE[tname])


def rel_rms(y, yref):
    y = np.asarray(y, np.float64)
    yref = np.asarray(yref, np.float64)
    den = np.sqrt(np.sum(yref * yref))
    return float(np.sqrt(np.sum((y - yref) ** 2)) / (den if den > 0 else 1.0))
