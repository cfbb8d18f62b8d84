"""kissfft_b200 -- Python plumbing over the C-ABI libraries libkissfft-<type>.so (sm_100a).

The product is the shared library (kissfft_b200/lib/libkissfft-{float,double,int16_t,int32_t}.so, built by
kissfft_b200.build from csrc/): it exports the reference's kiss_fft* API plus the batched device-pointer
entry points of include/kiss_fft_cuda.h.  This module only binds those C symbols with ctypes so that tests,
bench.py and multi-GPU drivers can call them on torch tensors; it contains no transform code and has no CPU
fallback -- if the library is missing, loading raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TYPES = ("float", "double", "int16_t", "int32_t")
NP_DTYPE = {"float": np.float32, "double": np.float64, "int16_t": np.int16, "int32_t": np.int32}

# every symbol include/*.h declares (the drop-in boundary)
API_SYMBOLS = (
    "kiss_fft_alloc", "kiss_fft", "kiss_fft_stride", "kiss_fft_cleanup", "kiss_fft_next_fast_size",
    "kiss_fftr_alloc", "kiss_fftr", "kiss_fftri",
    "kiss_fftnd_alloc", "kiss_fftnd",
    "kiss_fftndr_alloc", "kiss_fftndr", "kiss_fftndri",
    "kfc_fft", "kfc_ifft", "kfc_cleanup",
    "kiss_fft_batch_dev", "kiss_fftr_batch_dev", "kiss_fftri_batch_dev", "kiss_fftnd_dev", "kiss_fft_axis_pass_dev", "kiss_fft_planes_pass_dev", "kiss_fft_planes_pass_peers_dev",
    "kiss_fft_planes_pass_peers2_dev",
    "kiss_fftnd_mgpu_get_id", "kiss_fftnd_mgpu_alloc", "kiss_fftnd_mgpu_exec", "kiss_fftnd_mgpu_free", "kiss_fftnd_mgpu_local_in_elems",
    "kiss_fftnd_mgpu_local_out_elems", "kiss_fftnd_mgpu_uses_p2p", "kiss_fftnd_mgpu_chunks", "kiss_fftnd_mgpu_a2a_bytes",
    "kiss_fftnd_mgpu_last_error", "kiss_fftnd_mgpu_tune", "kiss_fftnd_mgpu_knob", "kiss_fftnd_mgpu_trace", "kiss_fftnd_mgpu_debug_chunks",
    "kiss_fftndr_dev", "kiss_fftndri_dev", "kiss_fft_batch", "kiss_fftr_batch", "kiss_fftri_batch",
    "kiss_fft_cuda_last_error", "kiss_fft_cuda_launch_count", "kiss_fft_cuda_plan_kind", "kiss_fft_cuda_scalar_bytes",
    "kiss_fft_cuda_is_fixed_point", "kiss_fft_cuda_force_generic", "kiss_fft_cuda_set_grid_limit", "kiss_fft_cuda_debug_chunks",
)


# float / double builds only (the reference's fast FIR "won't work for fixed point", kiss_fastfir.c:152)
FASTCONV_SYMBOLS = ("kiss_fastconv_alloc", "kiss_fastconvr_alloc", "kiss_fastconv_free", "kiss_fastconv_block_advance",
                    "kiss_fastconv_nfft", "kiss_fastconv_dev", "kiss_fastconvr_dev")


def lib_path(tname):
    return os.path.join(HERE, "lib", "libkissfft-%s.so" % tname)


class KissFFTError(RuntimeError):
    pass


def _ptr(x):
    """device/host address of a torch tensor, numpy array or int."""
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    return ctypes.c_void_p(x.data_ptr())   # torch.Tensor


class KissFFT:
    """One datatype build of the library.  Methods mirror the C entry points one to one."""

    def __init__(self, tname="float"):
        if tname not in TYPES:
            raise ValueError("datatype must be one of %s" % (TYPES,))
        path = lib_path(tname)
        if not os.path.exists(path):
            raise KissFFTError("%s is missing -- build it with `python -m kissfft_b200.build` (no CPU fallback exists)" % path)
        self.tname = tname
        self.dtype = NP_DTYPE[tname]
        self.path = path
        self.lib = L = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        vp, ci, sz, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_longlong
        L.kiss_fft_alloc.restype = vp
        L.kiss_fft_alloc.argtypes = [ci, ci, vp, ctypes.POINTER(sz)]
        L.kiss_fftr_alloc.restype = vp
        L.kiss_fftr_alloc.argtypes = [ci, ci, vp, ctypes.POINTER(sz)]
        L.kiss_fftnd_alloc.restype = vp
        L.kiss_fftnd_alloc.argtypes = [ctypes.POINTER(ci), ci, ci, vp, ctypes.POINTER(sz)]
        L.kiss_fftndr_alloc.restype = vp
        L.kiss_fftndr_alloc.argtypes = [ctypes.POINTER(ci), ci, ci, vp, ctypes.POINTER(sz)]
        for name in ("kiss_fft", "kiss_fftr", "kiss_fftri", "kiss_fftnd", "kiss_fftndr", "kiss_fftndri"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, vp, vp]
        L.kiss_fft_stride.restype = None
        L.kiss_fft_stride.argtypes = [vp, vp, vp, ci]
        L.kiss_fft_cleanup.restype = None
        for name in ("kfc_fft", "kfc_ifft"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [ci, vp, vp]
        L.kfc_cleanup.restype = None
        L.kiss_fft_next_fast_size.argtypes = [ci]
        L.kiss_fft_batch_dev.argtypes = [vp, vp, vp, sz, sz, sz, ci, vp]
        L.kiss_fftr_batch_dev.argtypes = [vp, vp, vp, sz, sz, sz, vp]
        L.kiss_fftri_batch_dev.argtypes = [vp, vp, vp, sz, sz, sz, vp]
        L.kiss_fftnd_dev.argtypes = [vp, vp, vp, vp, vp]
        L.kiss_fft_axis_pass_dev.argtypes = [vp, vp, vp, sz, sz, vp]
        L.kiss_fft_planes_pass_dev.argtypes = [vp, vp, vp, sz, sz, sz, sz, sz, vp]
        L.kiss_fft_planes_pass_peers_dev.argtypes = [vp, vp, ctypes.POINTER(vp), ci, sz, sz, sz, sz, sz, vp]
        L.kiss_fft_planes_pass_peers2_dev.argtypes = [vp, vp, ctypes.POINTER(vp), ci, sz, sz, sz, sz, sz, sz, sz, ci, vp]
        L.kiss_fftnd_mgpu_get_id.argtypes = [vp]
        L.kiss_fftnd_mgpu_alloc.restype = vp
        L.kiss_fftnd_mgpu_alloc.argtypes = [ctypes.POINTER(ci), ci, ci, ci, ci, vp, ctypes.c_uint]
        L.kiss_fftnd_mgpu_exec.argtypes = [vp, vp, vp, vp]
        L.kiss_fftnd_mgpu_free.argtypes = [vp]
        L.kiss_fftnd_mgpu_free.restype = None
        for name in ("kiss_fftnd_mgpu_local_in_elems", "kiss_fftnd_mgpu_local_out_elems", "kiss_fftnd_mgpu_a2a_bytes"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = sz
        for name in ("kiss_fftnd_mgpu_uses_p2p", "kiss_fftnd_mgpu_chunks"):
            getattr(L, name).argtypes = [vp]
        L.kiss_fftnd_mgpu_last_error.restype = ctypes.c_char_p
        L.kiss_fftnd_mgpu_tune.argtypes = [vp, ctypes.POINTER(ci), ci]
        L.kiss_fftnd_mgpu_knob.argtypes = [vp, ci]
        L.kiss_fftndr_dev.argtypes = [vp, vp, vp, vp]
        L.kiss_fftndri_dev.argtypes = [vp, vp, vp, vp]
        L.kiss_fft_batch.argtypes = [vp, vp, vp, sz]
        L.kiss_fftr_batch.argtypes = [vp, vp, vp, sz]
        L.kiss_fftri_batch.argtypes = [vp, vp, vp, sz]
        L.kiss_fft_cuda_last_error.restype = ctypes.c_char_p
        L.kiss_fft_cuda_launch_count.restype = ll
        L.kiss_fft_cuda_plan_kind.argtypes = [ci]
        L.kiss_fft_cuda_force_generic.argtypes = [ci]
        L.kiss_fft_cuda_force_generic.restype = None
        L.kiss_fft_cuda_set_grid_limit.argtypes = [ci]
        L.kiss_fft_cuda_set_grid_limit.restype = None
        if tname in ("float", "double"):
            L.kiss_fastconv_alloc.restype = vp
            L.kiss_fastconv_alloc.argtypes = [vp, sz, ctypes.POINTER(sz)]
            L.kiss_fastconv_free.argtypes = [vp]
            L.kiss_fastconv_free.restype = None
            L.kiss_fastconv_block_advance.argtypes = [vp]
            L.kiss_fastconv_block_advance.restype = sz
            L.kiss_fastconv_nfft.argtypes = [vp]
            L.kiss_fastconv_nfft.restype = sz
            L.kiss_fastconv_dev.argtypes = [vp, vp, vp, sz, ctypes.POINTER(sz), vp]
            L.kiss_fastconvr_alloc.restype = vp
            L.kiss_fastconvr_alloc.argtypes = [vp, sz, ctypes.POINTER(sz)]
            L.kiss_fastconvr_dev.argtypes = [vp, vp, vp, sz, ctypes.POINTER(sz), vp]
        self._libc = ctypes.CDLL(None)
        self._libc.free.argtypes = [vp]
        if L.kiss_fft_cuda_scalar_bytes() != np.dtype(self.dtype).itemsize:
            raise KissFFTError("%s was built for a different datatype" % path)

    # ---- plans (cfg objects are plain malloc blocks: release with free(), kiss_fft.h:138) ----
    def alloc(self, nfft, inverse=False):
        cfg = self.lib.kiss_fft_alloc(int(nfft), int(bool(inverse)), None, None)
        if not cfg:
            raise KissFFTError("kiss_fft_alloc(%d) returned NULL" % nfft)
        return cfg

    def allocr(self, nfft, inverse=False):
        cfg = self.lib.kiss_fftr_alloc(int(nfft), int(bool(inverse)), None, None)
        if not cfg:
            raise KissFFTError("kiss_fftr_alloc(%d) returned NULL" % nfft)
        return cfg

    def allocnd(self, dims, inverse=False):
        arr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
        cfg = self.lib.kiss_fftnd_alloc(arr, len(dims), int(bool(inverse)), None, None)
        if not cfg:
            raise KissFFTError("kiss_fftnd_alloc(%s) returned NULL" % (dims,))
        return cfg

    def allocndr(self, dims, inverse=False):
        arr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
        cfg = self.lib.kiss_fftndr_alloc(arr, len(dims), int(bool(inverse)), None, None)
        if not cfg:
            raise KissFFTError("kiss_fftndr_alloc(%s) returned NULL" % (dims,))
        return cfg

    def free(self, cfg):
        self._libc.free(ctypes.c_void_p(cfg))

    def cleanup(self):
        self.lib.kiss_fft_cleanup()

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.kiss_fft_cuda_last_error()
            raise KissFFTError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else ""))

    # ---- device-pointer batched calls (tensors or raw addresses) ----
    def fft_batch_dev(self, cfg, d_in, d_out, howmany, in_dist, out_dist, in_stride=1, stream=0):
        self._check(self.lib.kiss_fft_batch_dev(cfg, _ptr(d_in), _ptr(d_out), howmany, in_dist, out_dist, in_stride,
                                                ctypes.c_void_p(stream)), "kiss_fft_batch_dev")

    def fftr_batch_dev(self, cfg, d_time, d_freq, howmany, time_dist, freq_dist, stream=0):
        self._check(self.lib.kiss_fftr_batch_dev(cfg, _ptr(d_time), _ptr(d_freq), howmany, time_dist, freq_dist,
                                                 ctypes.c_void_p(stream)), "kiss_fftr_batch_dev")

    def fftri_batch_dev(self, cfg, d_freq, d_time, howmany, freq_dist, time_dist, stream=0):
        self._check(self.lib.kiss_fftri_batch_dev(cfg, _ptr(d_freq), _ptr(d_time), howmany, freq_dist, time_dist,
                                                  ctypes.c_void_p(stream)), "kiss_fftri_batch_dev")

    def fftnd_dev(self, cfg, d_in, d_out, d_work=None, stream=0):
        self._check(self.lib.kiss_fftnd_dev(cfg, _ptr(d_in), _ptr(d_out), _ptr(d_work), ctypes.c_void_p(stream)),
                    "kiss_fftnd_dev")

    def axis_pass_dev(self, cfg, d_in, d_out, ncols, col_stride, stream=0):
        self._check(self.lib.kiss_fft_axis_pass_dev(cfg, _ptr(d_in), _ptr(d_out), ncols, col_stride,
                                                    ctypes.c_void_p(stream)), "kiss_fft_axis_pass_dev")

    def planes_pass_dev(self, cfg, d_in, d_out, nplanes, ncols, col_stride, in_plane_dist, out_plane_dist, stream=0):
        self._check(self.lib.kiss_fft_planes_pass_dev(cfg, _ptr(d_in), _ptr(d_out), nplanes, ncols, col_stride, in_plane_dist,
                                                      out_plane_dist, ctypes.c_void_p(stream)), "kiss_fft_planes_pass_dev")

    def planes_pass_peers_dev(self, cfg, d_in, peer_ptrs, nplanes, cols_per_peer, col_stride, in_plane_dist, out_plane_dist,
                              stream=0):
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self.lib.kiss_fft_planes_pass_peers_dev(cfg, _ptr(d_in), arr, len(peer_ptrs), nplanes, cols_per_peer,
                                                            col_stride, in_plane_dist, out_plane_dist,
                                                            ctypes.c_void_p(stream)), "kiss_fft_planes_pass_peers_dev")

    # ---- kiss_fftnd over several GPUs (one process per GPU), include/kiss_fft_cuda.h ----
    MGPU_P2P = 1
    MGPU_REFERENCE_ORDER = 2

    def mgpu_get_id(self):
        """rank 0: the 128-byte rendezvous id to hand to every rank's mgpu_alloc"""
        buf = ctypes.create_string_buffer(128)
        rc = self.lib.kiss_fftnd_mgpu_get_id(buf)
        if rc != 0:
            raise KissFFTError("kiss_fftnd_mgpu_get_id failed (%d): %s" % (rc, (self.lib.kiss_fftnd_mgpu_last_error() or b"").decode()))
        return buf.raw

    def mgpu_alloc(self, dims, rank, nranks, ident=None, inverse=False, flags=0):
        arr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
        idbuf = ctypes.create_string_buffer(ident, 128) if ident is not None else None
        cfg = self.lib.kiss_fftnd_mgpu_alloc(arr, len(dims), int(bool(inverse)), int(rank), int(nranks), idbuf, int(flags))
        if not cfg:
            raise KissFFTError("kiss_fftnd_mgpu_alloc failed: %s" % (self.lib.kiss_fftnd_mgpu_last_error() or b"").decode())
        return cfg

    def mgpu_exec(self, cfg, d_in, d_out, stream=0):
        rc = self.lib.kiss_fftnd_mgpu_exec(cfg, _ptr(d_in), _ptr(d_out), ctypes.c_void_p(stream))
        if rc != 0:
            raise KissFFTError("kiss_fftnd_mgpu_exec failed (%d): %s" % (rc, (self.lib.kiss_fftnd_mgpu_last_error() or b"").decode()))

    def mgpu_free(self, cfg):
        self.lib.kiss_fftnd_mgpu_free(cfg)

    def mgpu_info(self, cfg):
        L = self.lib
        return {"in_elems": int(L.kiss_fftnd_mgpu_local_in_elems(cfg)), "out_elems": int(L.kiss_fftnd_mgpu_local_out_elems(cfg)),
                "p2p": bool(L.kiss_fftnd_mgpu_uses_p2p(cfg)), "chunks": int(L.kiss_fftnd_mgpu_chunks(cfg)),
                "a2a_bytes": int(L.kiss_fftnd_mgpu_a2a_bytes(cfg)), "pchunks": int(L.kiss_fftnd_mgpu_knob(cfg, 1)),
                "b_ctas": int(L.kiss_fftnd_mgpu_knob(cfg, 2)), "b_prio": int(L.kiss_fftnd_mgpu_knob(cfg, 3)),
                "ac_reserve": int(L.kiss_fftnd_mgpu_knob(cfg, 4)), "link_sms": int(L.kiss_fftnd_mgpu_knob(cfg, 6)),
                "rest_sms": int(L.kiss_fftnd_mgpu_knob(cfg, 7))}

    def fftndr_dev(self, cfg, d_time, d_freq, stream=0):
        self._check(self.lib.kiss_fftndr_dev(cfg, _ptr(d_time), _ptr(d_freq), ctypes.c_void_p(stream)), "kiss_fftndr_dev")

    def fftndri_dev(self, cfg, d_freq, d_time, stream=0):
        self._check(self.lib.kiss_fftndri_dev(cfg, _ptr(d_freq), _ptr(d_time), ctypes.c_void_p(stream)), "kiss_fftndri_dev")

    # ---- host-pointer batched calls ----
    def fft_batch(self, cfg, h_in, h_out, howmany):
        self._check(self.lib.kiss_fft_batch(cfg, _ptr(h_in), _ptr(h_out), howmany), "kiss_fft_batch")

    def fftr_batch(self, cfg, h_time, h_freq, howmany):
        self._check(self.lib.kiss_fftr_batch(cfg, _ptr(h_time), _ptr(h_freq), howmany), "kiss_fftr_batch")

    def fftri_batch(self, cfg, h_freq, h_time, howmany):
        self._check(self.lib.kiss_fftri_batch(cfg, _ptr(h_freq), _ptr(h_time), howmany), "kiss_fftri_batch")

    # ---- the reference's own calls (host or device pointers) ----
    def fft(self, cfg, fin, fout):
        self.lib.kiss_fft(cfg, _ptr(fin), _ptr(fout))

    def fft_stride(self, cfg, fin, fout, stride):
        self.lib.kiss_fft_stride(cfg, _ptr(fin), _ptr(fout), stride)

    def fftr(self, cfg, timedata, freqdata):
        self.lib.kiss_fftr(cfg, _ptr(timedata), _ptr(freqdata))

    def fftri(self, cfg, freqdata, timedata):
        self.lib.kiss_fftri(cfg, _ptr(freqdata), _ptr(timedata))

    def fftnd(self, cfg, fin, fout):
        self.lib.kiss_fftnd(cfg, _ptr(fin), _ptr(fout))

    def fftndr(self, cfg, timedata, freqdata):
        self.lib.kiss_fftndr(cfg, _ptr(timedata), _ptr(freqdata))

    def fftndri(self, cfg, freqdata, timedata):
        self.lib.kiss_fftndri(cfg, _ptr(freqdata), _ptr(timedata))

    # ---- fused fast convolution (float / double) ----
    def fastconv_alloc(self, imp_resp, nfft=0):
        """imp_resp: (n, 2) array of the build's scalar type.  Returns (cfg, nfft, ngood)."""
        imp = np.ascontiguousarray(imp_resp, dtype=self.dtype)
        n = ctypes.c_size_t(int(nfft))
        cfg = self.lib.kiss_fastconv_alloc(_ptr(imp), imp.shape[0], ctypes.byref(n))
        if not cfg:
            raise KissFFTError("kiss_fastconv_alloc failed")
        return cfg, int(n.value), int(self.lib.kiss_fastconv_block_advance(cfg))

    def fastconv_dev(self, cfg, d_in, d_out, n, stream=0):
        done = ctypes.c_size_t(0)
        self._check(self.lib.kiss_fastconv_dev(cfg, _ptr(d_in), _ptr(d_out), n, ctypes.byref(done), ctypes.c_void_p(stream)),
                    "kiss_fastconv_dev")
        return int(done.value)

    def fastconvr_alloc(self, imp_resp, nfft=0):
        """real samples (the reference's REAL_FASTFIR build): imp_resp is a 1-D array.  Returns (cfg, nfft, ngood)."""
        imp = np.ascontiguousarray(imp_resp, dtype=self.dtype)
        n = ctypes.c_size_t(int(nfft))
        cfg = self.lib.kiss_fastconvr_alloc(_ptr(imp), imp.shape[0], ctypes.byref(n))
        if not cfg:
            raise KissFFTError("kiss_fastconvr_alloc failed")
        return cfg, int(n.value), int(self.lib.kiss_fastconv_block_advance(cfg))

    def fastconvr_dev(self, cfg, d_in, d_out, n, stream=0):
        done = ctypes.c_size_t(0)
        self._check(self.lib.kiss_fastconvr_dev(cfg, _ptr(d_in), _ptr(d_out), n, ctypes.byref(done), ctypes.c_void_p(stream)),
                    "kiss_fastconvr_dev")
        return int(done.value)

    def fastconv_free(self, cfg):
        self.lib.kiss_fastconv_free(cfg)

    # ---- introspection ----
    def launch_count(self):
        return int(self.lib.kiss_fft_cuda_launch_count())

    def plan_kind(self, nfft):
        return int(self.lib.kiss_fft_cuda_plan_kind(int(nfft)))

    def force_generic(self, on):
        self.lib.kiss_fft_cuda_force_generic(int(bool(on)))

    def set_grid_limit(self, max_ctas):
        self.lib.kiss_fft_cuda_set_grid_limit(int(max_ctas))

    def last_error(self):
        msg = self.lib.kiss_fft_cuda_last_error()
        return msg.decode() if msg else ""

    def next_fast_size(self, n):
        return int(self.lib.kiss_fft_next_fast_size(int(n)))


_LIBS = {}


def get(tname="float"):
    """process-wide KissFFT instance for a datatype."""
    if tname not in _LIBS:
        _LIBS[tname] = KissFFT(tname)
    return _LIBS[tname]
