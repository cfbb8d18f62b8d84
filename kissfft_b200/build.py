"""Builds libkissfft-<type>.so (one per datatype, like the reference's Makefile:89-137) for sm_100a, in-tree.

    python -m kissfft_b200.build            # all four datatypes
    python -m kissfft_b200.build float      # a subset

Host layer (csrc/kf_api.c) is compiled as C by gcc, kernels + launchers (csrc/kf_launch.cu) by nvcc for
arch=compute_100a,code=sm_100a only; the two objects are linked into kissfft_b200/lib/libkissfft-<type>.so.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
HOSTCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

TYPES = ("float", "double", "int16_t", "int32_t")
TYPEFLAGS = {
    "float": ["-Dkiss_fft_scalar=float"],
    "double": ["-Dkiss_fft_scalar=double", "-DKF_IS_DOUBLE"],
    "int16_t": ["-DFIXED_POINT=16"],
    "int32_t": ["-DFIXED_POINT=32"],
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path(tname):
    return os.path.join(LIBDIR, "libkissfft-%s.so" % tname)


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(HERE, "..", "include", f) for f in sorted(os.listdir(os.path.join(HERE, "..", "include"))) if f.endswith(".h")]


def _stale(target, sources=None):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in (sources if sources is not None else _sources()))


def _cuda_sources():
    """what the kernel object depends on: everything but the C host layer"""
    return [s for s in _sources() if not s.endswith(".c")]


def _host_sources():
    return [s for s in _sources() if s.endswith((".c", "kf_internal.h")) or os.sep + "include" + os.sep in s]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))
    return r.stdout + r.stderr


def build_one(tname, force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    out = lib_path(tname)
    if not force and not _stale(out):
        return out
    tf = TYPEFLAGS[tname]
    o_cu = os.path.join(OBJDIR, "kf_launch-%s.o" % tname)
    o_c = os.path.join(OBJDIR, "kf_api-%s.o" % tname)
    o_m = os.path.join(OBJDIR, "kf_mgpu-%s.o" % tname)
    if force or _stale(o_cu, _cuda_sources()):
        log = _run([NVCC, "-std=c++20", "--expt-relaxed-constexpr", *ARCH, "-lineinfo", "-O3", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                    "-ccbin", HOSTCXX, "-Xptxas", "-v", "-DKISS_FFT_SHARED", *tf, "-c", os.path.join(CSRC, "kf_launch.cu"), "-o", o_cu])
        if verbose:
            print(log)
    for src, obj in (("kf_api.c", o_c), ("kf_mgpu.c", o_m)):
        if force or _stale(obj, _host_sources()):
            _run([HOSTCC, "-std=gnu11", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-DKISS_FFT_SHARED", *tf,
                  "-I", os.path.join(CUDA_HOME, "include"), "-c", os.path.join(CSRC, src), "-o", obj])
    _run([NVCC, "-shared", *ARCH, "-ccbin", HOSTCXX, o_cu, o_c, o_m, "-o", out, "-lpthread", "-lm", "-ldl"])
    return out


def build_all(types=TYPES, force=False, verbose=False):
    with cf.ThreadPoolExecutor(max_workers=len(types)) as ex:
        return list(ex.map(lambda t: build_one(t, force, verbose), types))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    force = "--force" in sys.argv
    verbose = "-v" in sys.argv
    for p in build_all(tuple(args) if args else TYPES, force, verbose):
        print(p)
