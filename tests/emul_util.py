"""ctypes wrapper of tests/emul (CPU emulator of the kernel bodies; test-only, see emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul")
FLAGS = {"float": "-Dkiss_fft_scalar=float", "double": "-Dkiss_fft_scalar=double -DKF_IS_DOUBLE", "int16_t": "-DFIXED_POINT=16",
         "int32_t": "-DFIXED_POINT=32"}
C2C, C2C_COL, R2C, C2R = range(4)


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    csrc = os.path.join(HERE, "..", "kissfft_b200", "csrc")
    deps = [os.path.join(EMUL, "emul.cpp"), os.path.join(EMUL, "experimental_plans.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(tname):
    lib = os.path.join(EMUL, "_build", "libemul-%s.so" % tname)
    if _stale(lib):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", *FLAGS[tname].split(),
                        os.path.join(EMUL, "emul.cpp"), "-o", lib], check=True)
    return lib


def build_all(types=("float", "double", "int16_t", "int32_t")):
    """compile the four emulator libraries concurrently (each takes about a minute of g++ time when stale)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(types)) as ex:
        return list(ex.map(build, types))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Emulator:
    def __init__(self, tname):
        self.lib = ctypes.CDLL(build(tname))
        vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
        self.lib.emul_fused.argtypes = [ci, ci, ci, vp, vp, ll, ll, ll, ll, vp, vp, ll]
        self.lib.emul_fourstep.argtypes = [ci, ci, ci, vp, vp, ll, ll, vp, vp, ll]
        self.lib.emul_colring.argtypes = [ci, ci, ci, vp, vp, ll, ll, ll, ll, ll, ll, vp, vp, ll]
        self.lib.emul_fused_experimental.argtypes = [ci, ci, ci, vp, vp, ll, ll, ll, ll, vp, vp, ll]
        self.lib.emul_generic.argtypes = [ci, ci, ci, vp, ci, vp, vp, ll, ll, ll, ll, vp, vp, ci, ci, ll]
        self.lib.emul_stage.argtypes = [ci, ci, ci, ci, ci, vp, vp, ll, ll, ll, ll, ci, ci, vp, ci, ll]
        self.lib.emul_realpass.argtypes = [ci, ci, vp, vp, ll, ll, ll, vp, ci, ll]
        if hasattr(self.lib, "emul_fastconv"):
            self.lib.emul_fastconv.argtypes = [ci, vp, vp, ll, ll, vp, vp, vp, ll]

    def plans(self):
        return [(self.lib.emul_plan_nfft(i), [m for m in range(4) if self.lib.emul_plan_has_mode(i, m)])
                for i in range(self.lib.emul_num_plans())]

    def experimental(self):
        """[(index, nfft, modes)] of the test-only plan variants (tests/emul/experimental_plans.h)"""
        return [(i, self.lib.emul_experimental_nfft(i), [m for m in range(4) if self.lib.emul_experimental_has_mode(i, m)])
                for i in range(self.lib.emul_num_experimental())]

    def fused(self, nfft, mode, inverse, inp, out, howmany, in_dist, out_dist, in_stride, tw, stw=None, nblocks=2, factors=None,
              experimental=None):
        """mirrors kf_launch.cu: the fused kernel takes the rows its alignment rules allow, the run-time kernel the rest"""
        if experimental is not None:
            done = self.lib.emul_fused_experimental(experimental, mode, int(inverse), _p(inp), _p(out), howmany, in_dist, out_dist,
                                                    in_stride, _p(tw), _p(stw), nblocks)
        else:
            done = self.lib.emul_fused(nfft, mode, int(inverse), _p(inp), _p(out), howmany, in_dist, out_dist, in_stride,
                                       _p(tw), _p(stw), nblocks)
        assert done >= 0, "no fused plan for nfft=%d mode=%d" % (nfft, mode)
        if done < howmany:
            assert factors is not None, "ragged tail needs the run-time kernel"
            esz = inp.dtype.itemsize * 2
            tail_in = ctypes.c_void_p(inp.ctypes.data + done * in_dist * esz)
            tail_out = ctypes.c_void_p(out.ctypes.data + done * out_dist * esz)
            fac = np.array([x for pm in factors for x in pm], np.int32)
            rc = self.lib.emul_generic(nfft, mode, int(inverse), _p(fac), len(factors), tail_in, tail_out, howmany - done,
                                       in_dist, out_dist, in_stride, _p(tw), _p(stw), 2, 32, 1)
            assert rc == 0
        return done

    def fourstep(self, n1, n2, inverse, inp, out, nrows, tw1, tw2, twbig, nblocks=3):
        """rows of length n1*n2 by the two four-step passes (mirrors kf_api.c:kf_exec_fourstep); False when the plans are missing"""
        work = np.zeros_like(inp)
        a = self.lib.emul_fourstep(n1, 0, int(inverse), _p(inp), _p(work), nrows, n2, _p(tw1), _p(twbig), nblocks)
        if a < 0:
            return False
        b = self.lib.emul_fourstep(n2, 1, int(inverse), _p(work), _p(out), nrows, n1, _p(tw2), None, nblocks)
        return b >= 0

    def colcol(self, nfft, inverse, inp, out, nplanes, ncols, tw, nblocks=3):
        """axis pass that keeps the layout: plane p, column c: inp[p][j][c] (j < nfft) -> out[p][k][c]; False without a plan"""
        return self.lib.emul_fourstep(nfft, 1, int(inverse), _p(inp), _p(out), nplanes, ncols, _p(tw), None, nblocks) >= 0

    def colring(self, nfft, mode, inverse, inp, out, nplanes, ncols, col_stride, in_pdist, out_pdist, out_dist, tw, twbig=None, nblocks=2):
        """column modes through the tensor-map ring plans; returns the emulator's code (< 0: no plan / not eligible)"""
        return self.lib.emul_colring(nfft, mode, int(inverse), _p(inp), _p(out), nplanes, ncols, col_stride, in_pdist, out_pdist,
                                     out_dist, _p(tw), _p(twbig), nblocks)

    def generic(self, nfft, mode, inverse, factors, inp, out, howmany, in_dist, out_dist, in_stride, tw, stw=None,
                tpc=2, nthreads=32, nblocks=2):
        fac = np.array([x for pm in factors for x in pm], np.int32)
        rc = self.lib.emul_generic(nfft, mode, int(inverse), _p(fac), len(factors), _p(inp), _p(out), howmany, in_dist,
                                   out_dist, in_stride, _p(tw), _p(stw), tpc, nthreads, nblocks)
        assert rc == 0


    def multipass(self, nfft, inverse, factors, inp, out, howmany, in_dist, out_dist, in_stride, tw):
        """mirrors kf_api.c:kf_exec_multipass for the complex modes: one stage per call, dense work buffers"""
        L = len(factors)
        w = [np.zeros((howmany, nfft, 2), inp.dtype), np.zeros((howmany, nfft, 2), inp.dtype)]
        cur, F = inp, 1
        Fs = []
        for p, m in factors:
            Fs.append(F)
            F *= p
        wi = 0
        for s in range(L - 1, -1, -1):
            first, last = int(s == L - 1), int(s == 0)
            dst = out if last else w[wi]
            p, m = factors[s]
            rc = self.lib.emul_stage(nfft, int(inverse), p, m, Fs[s], _p(cur), _p(dst), howmany, in_dist if first else nfft,
                                     out_dist if last else nfft, in_stride if first else 1, first, last, _p(tw), 64, 3)
            assert rc == 0
            cur = dst
            wi ^= 1

    def realpass(self, nc, post, inp, out, howmany, in_dist, out_dist, stw):
        assert self.lib.emul_realpass(nc, int(post), _p(inp), _p(out), howmany, in_dist, out_dist, _p(stw), 64, 3) == 0

    def fastconv(self, nfft, inp, out, nblocks, ngood, h, tw_f, tw_i, grid=2):
        rc = self.lib.emul_fastconv(nfft, _p(inp), _p(out), nblocks, ngood, _p(h), _p(tw_f), _p(tw_i), grid)
        assert rc == 0, "no fast-convolution plan for nfft=%d" % nfft
