#!/usr/bin/env python
"""bench.py -- measures the batched transform hot path on B200 (and the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json
configs[1]: 1-D real float nfft=4096, batch=32768, forward (kiss_fftr) then inverse (kiss_fftri) -- two kernel
launches per step.  Other BASELINE configs are selectable with --workload for investigation; they are parity
test cases, not the headline line.

Printed JSON (one line, rank 0):
  value        whole-job GFLOP/s with inputs resident in HBM (5*N*log2N per complex transform, 2.5*N*log2N per
               real transform), CUDA-event timed, max over ranks
  e2e          same metric through the host-pointer C-ABI calls (kiss_fftr_batch / kiss_fftri_batch ...) with
               pinned host buffers: H2D + kernels + D2H inside the timed region
  roofline     dominant kernel: algorithmic HBM bytes per launch / its mean launch time (CUDA events around every
               launch inside the timed region) vs the measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the compiled reference (oracle/_ref) -- or the oracle port when that is absent -- on the host
               cores, bounded sample of the same workload
  e2e_pageable the same host-pointer calls on plain malloc()ed (numpy) buffers -- what a drop-in caller passes
  configs      every other BASELINE.json config measured in the same run (kernel ms, GFLOP/s, roofline fraction,
               parity flag against the CPU reference on a sample); N == 1: configs[0], [2], [3] and kiss_fftnd 1024^3;
               N > 1: the slab-sharded 1024^3 transform (strong scaling; fused peer-store exchange and NCCL), with the
               all-to-all bytes against the NVLink roofline and the 1-GPU time measured on rank 0 in the same run
`--impl reference` times only the CPU reference arm and prints the same line shape (thread count = host cores,
set explicitly: torchrun exports OMP_NUM_THREADS=1).
Multi-GPU (torchrun, one rank per GPU): batches shard by rank with no communication => weak scaling.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched FFT GFLOP/s (5N*log2N)"

# name -> dict(kind, tname, nfft|dims, batch)
WORKLOADS = {
    "r2c4096": dict(kind="real", tname="float", nfft=4096, batch=32768,
                    desc="1-D real float R2C+C2R nfft=4096 batch=32768 (kiss_fftr/kiss_fftri) -- BASELINE configs[1]"),
    "c2c1024": dict(kind="c2c", tname="float", nfft=1024, batch=65536,
                    desc="1-D complex float C2C nfft=1024 batch=65536 -- BASELINE configs[0]"),
    "c2c1000": dict(kind="c2c", tname="float", nfft=1000, batch=100000, desc="mixed radix nfft=1000 float -- configs[2]"),
    "c2c1155": dict(kind="c2c", tname="float", nfft=1155, batch=100000, desc="mixed radix nfft=1155 float -- configs[2]"),
    "z2z1000": dict(kind="c2c", tname="double", nfft=1000, batch=100000, desc="mixed radix nfft=1000 double -- configs[2]"),
    "z2z1155": dict(kind="c2c", tname="double", nfft=1155, batch=100000, desc="mixed radix nfft=1155 double -- configs[2]"),
    "q15_2048": dict(kind="c2c", tname="int16_t", nfft=2048, batch=65536, desc="Q15 C2C nfft=2048 batch=65536 -- configs[3]"),
    "q31_2048": dict(kind="c2c", tname="int32_t", nfft=2048, batch=65536, desc="Q31 C2C nfft=2048 batch=65536 -- configs[3]"),
    "fftnd256": dict(kind="nd", tname="float", dims=(256, 256, 256), batch=1, desc="3-D complex float 256^3 kiss_fftnd"),
    "fftnd512": dict(kind="nd", tname="float", dims=(512, 512, 512), batch=1, desc="3-D complex float 512^3 kiss_fftnd"),
    "fftnd1024": dict(kind="nd", tname="float", dims=(1024, 1024, 1024), batch=1,
                      desc="3-D complex float 1024^3 kiss_fftnd (single GPU) -- configs[4]"),
    # slab-decomposed 3-D transform: ONE array split over the ranks (strong scaling), one all-to-all
    "slab1024": dict(kind="slab", tname="float", dims=(1024, 1024, 1024), batch=1, p2p=True,
                     desc="3-D complex float 1024^3 slab-sharded, fused peer-memory exchange -- configs[4]"),
    "slab1024nccl": dict(kind="slab", tname="float", dims=(1024, 1024, 1024), batch=1, p2p=False,
                         desc="3-D complex float 1024^3 slab-sharded, NCCL all_to_all_single -- configs[4]"),
    "slab512": dict(kind="slab", tname="float", dims=(512, 512, 512), batch=1, p2p=True,
                    desc="3-D complex float 512^3 slab-sharded, fused peer-memory exchange"),
}
DTYPE_NAME = {"float": "f32", "double": "f64", "int16_t": "q15", "int32_t": "q31"}
NP = {"float": np.float32, "double": np.float64, "int16_t": np.int16, "int32_t": np.int32}


def flops_per_step(w):
    if w["kind"] == "real":      # forward + inverse real transform
        n = w["nfft"]
        return 2 * 2.5 * n * math.log2(n) * w["batch"]
    if w["kind"] == "c2c":
        n = w["nfft"]
        return 5.0 * n * math.log2(n) * w["batch"]
    n = int(np.prod(w["dims"]))      # nd / slab: one 3-D transform
    return 5.0 * n * math.log2(n)


def esz(tname):
    return np.dtype(NP[tname]).itemsize


def algorithmic_bytes(w):
    """per launch of the dominant kernel: one read + one write of every element"""
    s = esz(w["tname"])
    if w["kind"] == "real":      # R2C launch: N scalars in, N/2+1 complex out (C2R is the mirror image)
        n = w["nfft"]
        return (n * s + (n // 2 + 1) * 2 * s) * w["batch"]
    if w["kind"] == "c2c":
        return 2 * w["nfft"] * 2 * s * w["batch"]
    return 2 * int(np.prod(w["dims"])) * 2 * s   # one axis pass (nd / slab)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic(name):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


# ---- clocks ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        sm, smax, reasons = [], [], set()
        with open(self.tmp.name) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ---- CPU reference arm -------------------------------------------------------------------------------------
_CPU_INPUT = {}


def cpu_reference(w, reps=5, budget_rows=None, detail=False):
    """times the reference's own CPU implementation (oracle/_ref when compiled here, else the oracle port) on a
    bounded sample of the workload with all host threads (batch-parallel harness; cfgs are read-only, reference
    README.md:217).  Returns (gflops, info dict)."""
    from oracle import loader
    tname = w["tname"]
    dtype = NP[tname]
    have_ref = loader.have_reference(tname)
    drv = loader.CpuDriver()
    # explicit thread count = the cores this process may run on; omp_get_max_threads() is NOT consulted because
    # torch.distributed.run exports OMP_NUM_THREADS=1 (the driver's `num_threads(n)` clause overrides the ICV)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    threads = cores if have_ref else 1
    s = esz(tname)
    if w["kind"] in ("real", "c2c"):
        n = w["nfft"]
        # the whole batch when one operand fits 1 GiB of host memory (true for every bench workload), else a prefix of it
        rows = budget_rows or (min(w["batch"], max(256, (1 << 30) // (n * 2 * s))) if have_ref else min(w["batch"], 2048))
        sub = dict(w, batch=rows)
        key = (tname, rows, n, w["kind"])
        if w["kind"] == "c2c":
            x = _CPU_INPUT.get(key)
            if x is None:
                x = _CPU_INPUT.setdefault(key, loader.random_input(tname, (rows, n), 1))
            out = np.empty_like(x)
            if have_ref:
                t = drv.run(loader.reference_lib_path(tname), loader.K_FFT, [n], 0, x, out, rows, n * 2 * s, n * 2 * s, threads, reps)
            else:
                o = loader.Oracle(tname)
                t0 = time.perf_counter()
                o.fft(x)
                t = time.perf_counter() - t0
        else:
            x = _CPU_INPUT.get(key)
            if x is None:
                x = _CPU_INPUT.setdefault(key, loader.random_input(tname, (rows, n), 1, complex_=False))
            X = np.empty((rows, n // 2 + 1, 2), dtype)
            y = np.empty_like(x)
            if have_ref:
                lp = loader.reference_lib_path(tname)
                t = drv.run(lp, loader.K_FFTR, [n], 0, x, X, rows, n * s, (n // 2 + 1) * 2 * s, threads, reps)
                t += drv.run(lp, loader.K_FFTRI, [n], 1, X, y, rows, (n // 2 + 1) * 2 * s, n * s, threads, reps)
            else:
                o = loader.Oracle(tname)
                t0 = time.perf_counter()
                o.fftri(o.fftr(x))
                t = time.perf_counter() - t0
        sample = "%d of %d transforms, %s" % (rows, w["batch"], "forward+inverse" if w["kind"] == "real" else "forward")
        gf = flops_per_step(sub) / t / 1e9
        other = {}
        if have_ref and detail:
            # (A) the reference's OpenMP build as shipped (a <= p-way split per transform,
            # kiss_fft.c:253-260), (B) the plain build on one thread; `value` is (C), the plain build with the batch
            # spread over all cores by the harness.  A and B run on a 2048-row prefix.
            r2 = min(rows, 2048)
            f2 = flops_per_step(dict(w, batch=r2))
            for key, lp in (("A_openmp_build_as_shipped", loader.reference_lib_path(tname, openmp=True)),
                            ("B_plain_build_1_thread", loader.reference_lib_path(tname))):
                if not os.path.exists(lp):
                    continue
                if w["kind"] == "c2c":
                    tt = drv.run(lp, loader.K_FFT, [n], 0, x, out, r2, n * 2 * s, n * 2 * s, 1, 2)
                else:
                    tt = drv.run(lp, loader.K_FFTR, [n], 0, x, X, r2, n * s, (n // 2 + 1) * 2 * s, 1, 2)
                    tt += drv.run(lp, loader.K_FFTRI, [n], 1, X, y, r2, (n // 2 + 1) * 2 * s, n * s, 1, 2)
                other[key] = round(f2 / tt / 1e9, 3)
    else:
        dims = w["dims"]
        sdims = tuple(min(d, 128) for d in dims)       # the CPU needs minutes beyond 256^3; bounded sample
        x = loader.random_input(tname, sdims, 1)
        out = np.empty_like(x)
        if have_ref:
            t = drv.run(loader.reference_lib_path(tname), loader.K_FFTND, list(sdims), 0, x, out, 1, 0, 0, 1, 1)
        else:
            t0 = time.perf_counter()
            loader.Oracle(tname).fftnd(x)
            t = time.perf_counter() - t0
        threads = 1
        sample = "%s grid instead of %s (kiss_fftnd is single-threaded)" % ("x".join(map(str, sdims)), "x".join(map(str, dims)))
        gf = flops_per_step(dict(w, dims=sdims)) / t / 1e9
    info = {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": "reference" if have_ref else "port", "sample": sample,
            "seconds": t, "host_cores": cores}
    if w["kind"] in ("real", "c2c") and other:
        info["other_modes_gflops"] = other
    return gf, info


# ---- GPU arm ----------------------------------------------------------------------------------------------
def _tdt(tname):
    import torch
    return {"float": torch.float32, "double": torch.float64, "int16_t": torch.int16, "int32_t": torch.int32}[tname]


def config_dict(w, name, world):
    """the `config` object: identical for the repo arm and the reference arm of one workload"""
    strong = w["kind"] == "slab"
    return {"workload": w["desc"], "name": name, "batch_per_gpu": w["batch"],
            "flop_convention": "5*N*log2N per complex transform, 2.5*N*log2N per real transform",
            "l2": "inputs larger than L2 (no flush needed)" if algorithmic_bytes(w) > 300e6 else "working set may fit L2",
            "parallelism": ("slabs of d0/G planes per GPU, one all-to-all (%s)" % ("fused peer-memory stores" if w.get("p2p") else "NCCL")
                            if strong else "batch sharded across the GPUs, no communication")}


class Synth:
    """seeded synthetic inputs generated on the device (SURVEY 8d value distributions)"""

    def __init__(self, tname, seed):
        import torch
        self.torch, self.tname, self.tdt = torch, tname, _tdt(tname)
        self.gen = torch.Generator(device="cuda")
        self.gen.manual_seed(seed)

    def __call__(self, shape):
        torch = self.torch
        if self.tname in ("float", "double"):
            return torch.rand(shape, generator=self.gen, device="cuda", dtype=self.tdt) * 2 - 1
        half = (32767 if self.tname == "int16_t" else 2147483647) // 2
        return torch.randint(-half, half + 1, shape, generator=self.gen, device="cuda", dtype=torch.int64).to(self.tdt)


def time_kernels(kernels, steps, warmup, barrier):
    """W warm-up steps, then `steps` timed steps with a CUDA event around every launch (on the launching stream).
    Returns (total ms of the timed region, mean ms per kernel)."""
    import torch
    for _ in range(max(warmup, 3)):
        for _, k in kernels:
            k()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(kernels) + 1)] for _ in range(steps)]
    barrier()
    for s in range(steps):
        ev[s][0].record()
        for j, (_, k) in enumerate(kernels):
            k()
            ev[s][j + 1].record()
    barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][-1])
    per = [float(np.mean([ev[s][j].elapsed_time(ev[s][j + 1]) for s in range(steps)])) for j in range(len(kernels))]
    return total_ms, per


def sampled_dft_sums(x, p0, dims, ks):
    """partial direct-DFT sums (float64, on the GPU) of the planes x = X_in[p0:p0+len(x)] for the output bins
    (ks[0][b], ks[1][b], ks[2][b]); SURVEY 8(d) parity sampling for arrays beyond the CPU oracle's reach."""
    import torch
    cd = torch.complex128
    nb = len(ks[0])
    W = [torch.exp(-2j * np.pi * (torch.arange(d, dtype=torch.float64)[:, None] * k[None, :].double()) / d).to(cd).cuda()
         for d, k in zip(dims, ks)]
    acc = torch.zeros(nb, dtype=cd, device="cuda")
    rest = int(np.prod(dims[1:]))
    step = max(1, (1 << 25) // rest)
    e = torch.zeros((), dtype=torch.float64, device="cuda")
    for q0 in range(0, x.shape[0], step):
        q1 = min(x.shape[0], q0 + step)
        xc = torch.view_as_complex(x[q0:q1].contiguous()).to(cd)
        e += (xc.real ** 2 + xc.imag ** 2).sum()
        t = xc.reshape(-1, dims[-1]) @ W[-1]
        del xc
        for a in range(len(dims) - 2, 0, -1):
            t = (t.reshape(-1, dims[a], nb) * W[a][None]).sum(1)
        acc += (t.reshape(q1 - q0, nb) * W[0][p0 + q0:p0 + q1]).sum(0)
    return acc, e


def energy(t):
    import torch
    e = torch.zeros((), dtype=torch.float64, device="cuda")
    step = max(1, (1 << 26) // max(1, t[0].numel()))
    for q0 in range(0, t.shape[0], step):
        e += (t[q0:q0 + step].double() ** 2).sum()
    return e


def parity_1d(w, lib, d_in, d_out, rows=64):
    """GPU output rows vs the CPU checker (compiled reference when present, else the oracle port) on a sample of the
    batch; the same call is the config's CPU baseline sample.  Returns (flag string, ok)."""
    from oracle import loader
    tname = w["tname"]
    b = d_in.shape[0]
    idx = np.unique(np.linspace(0, b - 1, rows).astype(np.int64))
    xin = d_in[idx].cpu().numpy()
    got = d_out[idx].cpu().numpy()
    o = loader.Reference(tname) if loader.have_reference(tname) else loader.Oracle(tname)
    want = o.fft(xin)
    if tname in ("float", "double"):
        err = loader.rel_rms(got, want)
        tol = (1e-6 if tname == "float" else 1e-14) * math.log2(w["nfft"])
        return "rel-rms %.2e <= %.1e on %d sampled rows" % (err, tol, len(idx)), bool(err <= tol)
    nbad = int(np.count_nonzero(got != want))
    return "bit-exact on %d sampled rows" % len(idx) if nbad == 0 else "%d scalars differ" % nbad, nbad == 0


def measure_config(name, args, barrier):
    """one BASELINE config other than the headline: device-resident kernel time, roofline fraction, parity flag"""
    import torch
    import kissfft_b200
    w = WORKLOADS[name]
    lib = kissfft_b200.get(w["tname"])
    stream = torch.cuda.current_stream().cuda_stream
    synth = Synth(w["tname"], 4321)
    peak, _ = measured_peak()
    out = {"workload": w["desc"], "dtype": DTYPE_NAME[w["tname"]]}
    if w["kind"] == "c2c":
        n, b = w["nfft"], w["batch"]
        d_x = synth((b, n, 2))
        d_X = torch.empty_like(d_x)
        cf = lib.alloc(n, False)
        total, per = time_kernels([("c2c", lambda: lib.fft_batch_dev(cf, d_x, d_X, b, n, n, 1, stream))], args.steps, args.warmup, barrier)
        out["parity"], out["parity_ok"] = parity_1d(w, lib, d_x, d_X)
        launches = 1
        abytes = algorithmic_bytes(w)
        lib.free(cf)
    else:                                   # nd
        dims = w["dims"]
        d_x = synth(tuple(dims) + (2,))
        d_X = torch.empty_like(d_x)
        cf = lib.allocnd(dims, False)
        l0 = lib.launch_count()
        lib.fftnd_dev(cf, d_x, d_X, None, stream)
        launches = lib.launch_count() - l0
        total, per = time_kernels([("fftnd", lambda: lib.fftnd_dev(cf, d_x, d_X, None, stream))], args.steps, args.warmup, barrier)
        g = torch.Generator(device="cpu")
        g.manual_seed(7)
        ks = [torch.randint(0, d, (32,), generator=g) for d in dims]
        acc, ein = sampled_dft_sums(d_x, 0, dims, ks)
        got = torch.view_as_complex(d_X[tuple(k.cuda() for k in ks)].contiguous()).to(torch.complex128)
        err = float((got - acc).abs().pow(2).sum().sqrt() / acc.abs().pow(2).sum().sqrt())
        pars = abs(float(energy(d_X)) / (float(np.prod(dims)) * float(ein)) - 1.0)
        tol = 1e-6 * math.log2(float(np.prod(dims)))
        out["parity"] = "32 sampled bins vs float64 DFT sums rel-rms %.2e, Parseval defect %.1e (<= %.1e)" % (err, pars, tol)
        out["parity_ok"] = bool(err <= tol and pars <= tol)
        abytes = algorithmic_bytes(w) * len(dims)
        lib.free(cf)
    ms = total / args.steps
    out.update({"ms": ms, "kernel_launches_per_step": int(launches), "gflops": flops_per_step(w) / (ms * 1e-3) / 1e9,
                "algorithmic_bytes": abytes, "achieved_GBps": abytes / (ms * 1e-3) / 1e9})
    out["frac"] = out["achieved_GBps"] / peak
    out["frac_of_nominal_8000"] = out["achieved_GBps"] / 8000.0
    if known_traffic(name):
        out["traffic"] = known_traffic(name)      # static: dram bytes per launch of the committed ncu capture (profiles/traffic.json)
    if w["tname"] in ("int16_t", "int32_t"):
        # fixed point is integer-pipe bound -- report the HBM fraction AND the integer-op rate
        out["int_gop_per_s_5NlogN"] = out["gflops"]
        out["bound"] = "integer issue slots, not HBM (ncu: profiles/r02/ncu_all_kernels.txt; exact C_FIXDIV/sround arithmetic fixes the instruction count)"
    del d_x, d_X
    torch.cuda.empty_cache()
    return out


def measure_slab(name, args, dist, rank, world, t1_ms):
    """slab-sharded 3-D transform (strong scaling) on all ranks, fenced per-step timing like everything else; parity by
    SURVEY 8(d) sampling: 32 bins of the distributed output vs float64 direct DFT sums all-reduced over the slabs"""
    import torch
    from kissfft_b200.slab import MgpuFFT3D
    w = WORKLOADS[name]
    dims = w["dims"]
    plan = MgpuFFT3D(dims, tname=w["tname"], p2p=w.get("p2p", False))     # the library's C-ABI (kiss_fftnd_mgpu_*)
    g = plan.geo
    sx, sout = plan.alloc()
    synth = Synth(w["tname"], 999 + rank)
    x0 = synth(tuple(sx.shape))
    sx.copy_(x0)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    plan.forward(sx, sout, stream)
    barrier()
    # parity on the result of THIS call (the timed loop below re-transforms sx in place, like the r01 bench did)
    gcpu = torch.Generator(device="cpu")
    gcpu.manual_seed(11)
    ks = [torch.randint(0, d, (32,), generator=gcpu) for d in dims]
    acc, ein = sampled_dft_sums(x0, g.plane_range()[0], dims, ks)
    buf = torch.view_as_real(acc).contiguous()
    dist.all_reduce(buf)
    tot = torch.stack([ein, energy(sout)])
    dist.all_reduce(tot)
    acc = torch.view_as_complex(buf)
    c0, c1 = g.col_range()
    mine = (ks[2] >= c0) & (ks[2] < c1)
    num = torch.zeros(2, dtype=torch.float64, device="cuda")
    if bool(mine.any()):
        sel = mine.nonzero().flatten()
        got = torch.view_as_complex(sout[(ks[2][sel] - c0).cuda(), ks[1][sel].cuda(), ks[0][sel].cuda()].contiguous()).to(torch.complex128)
        ref = acc[sel.cuda()]
        num[0] = (got - ref).abs().pow(2).sum()
        num[1] = ref.abs().pow(2).sum()
    dist.all_reduce(num)
    err = float((num[0] / num[1]).sqrt())
    pars = abs(float(tot[1]) / (float(np.prod(dims)) * float(tot[0])) - 1.0)
    tol = 1e-6 * math.log2(float(np.prod(dims)))
    del x0
    total, per = time_kernels([("slab3d", lambda: plan.forward(sx, sout, stream))], args.steps, args.warmup, barrier)
    t = torch.tensor([total / args.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    a2a = g.a2a_bytes_per_rank(2 * esz(w["tname"]))
    peak, _ = measured_peak()
    local_bytes = algorithmic_bytes(w) * 3 // world
    out = {"workload": w["desc"], "dtype": DTYPE_NAME[w["tname"]], "ms": ms, "scaling": "strong",
           "api": "kiss_fftnd_mgpu_exec (C-ABI)", "exchange": "peer stores over NVLink (CUDA IPC)" if plan.info["p2p"] else "NCCL grouped send/recv",
           "pipeline_chunks": plan.info["chunks"], "pipeline_plane_groups": plan.info.get("pchunks"),
           "link_bound_cta_cap": plan.info.get("b_ctas"), "link_stream_priority": plan.info.get("b_prio"),
           "sm_partition": {"link_bound_sms": plan.info.get("link_sms"), "hbm_bound_sms": plan.info.get("rest_sms"),
                            "how": "CUDA green contexts"} if plan.info.get("link_sms") else None,
           "gflops": flops_per_step(w) / (ms * 1e-3) / 1e9,
           "parity": "32 sampled bins vs float64 DFT sums rel-rms %.2e, Parseval defect %.1e (<= %.1e)" % (err, pars, tol),
           "parity_ok": bool(err <= tol and pars <= tol),
           "a2a_bytes_sent_per_gpu": int(a2a), "a2a_ms_at_900GBps": a2a / 900e9 * 1e3,
           "local_hbm_ms_at_measured_peak": local_bytes / (peak * 1e9) * 1e3,
           "nvlink_frac_if_exchange_took_whole_step": a2a / (ms * 1e-3) / 900e9,
           "step_vs_bound": max(a2a / 900e9, local_bytes / (peak * 1e9)) / (ms * 1e-3)}
    if t1_ms:
        out["t1_ms_single_gpu_same_run"] = t1_ms
        out["strong_scaling_efficiency"] = t1_ms / (world * ms)
    plan.close()
    del plan, sx, sout
    torch.cuda.empty_cache()
    return out


def run_ours(args, w, rank, world, local_rank):
    import torch
    import kissfft_b200

    torch.cuda.set_device(local_rank)
    dist = None
    # (the image exports NCCL_DEBUG=VERSION, which makes NCCL print its banner on stdout: main() has moved fd 1 to stderr
    # for the duration of the run, and the variable is dropped so the banner does not appear at all)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = kissfft_b200.get(w["tname"])
    tdt = _tdt(w["tname"])
    stream = torch.cuda.current_stream().cuda_stream
    synth = Synth(w["tname"], 1234 + rank)

    kernels = []          # (label, callable) executed per step, each one kernel launch
    e2e_step = e2e_pageable_step = None
    h2d = d2h = 0
    if w["kind"] == "real":
        n, b = w["nfft"], w["batch"]
        nb = n // 2 + 1
        d_x = synth((b, n))
        d_X = torch.empty((b, nb, 2), device="cuda", dtype=tdt)
        d_y = torch.empty_like(d_x)
        cf, ci = lib.allocr(n, False), lib.allocr(n, True)
        kernels = [("r2c", lambda: lib.fftr_batch_dev(cf, d_x, d_X, b, n, nb, stream)),
                   ("c2r", lambda: lib.fftri_batch_dev(ci, d_X, d_y, b, nb, n, stream))]
        h_x = torch.empty((b, n), dtype=tdt).pin_memory()
        h_x.copy_(d_x)
        h_X = torch.empty((b, nb, 2), dtype=tdt).pin_memory()
        h_y = torch.empty((b, n), dtype=tdt).pin_memory()
        # plain malloc()ed buffers, what an unmodified caller of the reference passes (test/benchkiss.c:76-79)
        p_x = np.array(h_x.numpy(), copy=True)
        p_X = np.empty((b, nb, 2), NP[w["tname"]])
        p_y = np.empty((b, n), NP[w["tname"]])

        def e2e_step():
            lib.fftr_batch(cf, h_x, h_X, b)
            lib.fftri_batch(ci, h_X, h_y, b)

        def e2e_pageable_step():
            lib.fftr_batch(cf, p_x, p_X, b)
            lib.fftri_batch(ci, p_X, p_y, b)
        h2d = h_x.numel() * h_x.element_size() + h_X.numel() * h_X.element_size()
        d2h = h_X.numel() * h_X.element_size() + h_y.numel() * h_y.element_size()
    elif w["kind"] == "c2c":
        n, b = w["nfft"], w["batch"]
        d_x = synth((b, n, 2))
        d_X = torch.empty_like(d_x)
        cf = lib.alloc(n, False)
        kernels = [("c2c", lambda: lib.fft_batch_dev(cf, d_x, d_X, b, n, n, 1, stream))]
        h_x = torch.empty((b, n, 2), dtype=tdt).pin_memory()
        h_x.copy_(d_x)
        h_X = torch.empty((b, n, 2), dtype=tdt).pin_memory()
        p_x = np.array(h_x.numpy(), copy=True)
        p_X = np.empty((b, n, 2), NP[w["tname"]])

        def e2e_step():
            lib.fft_batch(cf, h_x, h_X, b)

        def e2e_pageable_step():
            lib.fft_batch(cf, p_x, p_X, b)
        h2d = d2h = h_x.numel() * h_x.element_size()
    elif w["kind"] == "slab":
        from kissfft_b200.slab import MgpuFFT3D
        plan = MgpuFFT3D(w["dims"], tname=w["tname"], p2p=w.get("p2p", False))
        sx, sout = plan.alloc()
        sx.copy_(synth(tuple(sx.shape)))
        kernels = [("slab3d", lambda: plan.forward(sx, sout, stream))]
    else:
        dims = w["dims"]
        d_x = synth(tuple(dims) + (2,))
        d_X = torch.empty_like(d_x)
        cf = lib.allocnd(dims, False)
        kernels = [("fftnd", lambda: lib.fftnd_dev(cf, d_x, d_X, None, stream))]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled by nvidia-smi (100 ms period) from the warm-up, through the timed region, to the end of a
    # ~1 s continuation of the very same launch loop: the timed region alone (K steps of ~1 ms) is shorter than one
    # sampling period, so the continuation is what gives the "under load" median; it is not part of any timing.
    clocks = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):          # warm-up (also builds device tables, sets kernel attributes)
        for _, k in kernels:
            k()
    barrier()
    # the timed region: K steps, one CUDA event around every launch (on the launching stream)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(kernels) + 1)] for _ in range(args.steps)]
    barrier()
    l0 = lib.launch_count()
    for s in range(args.steps):
        ev[s][0].record()
        for j, (_, k) in enumerate(kernels):
            k()
            ev[s][j + 1].record()
    barrier()
    launches = lib.launch_count() - l0
    total_ms = ev[0][0].elapsed_time(ev[-1][-1])
    per_kernel_ms = [float(np.mean([ev[s][j].elapsed_time(ev[s][j + 1]) for s in range(args.steps)])) for j in range(len(kernels))]
    # continuation for the clock sampler: every rank runs the SAME number of extra steps (the slab workload contains
    # collectives, so a rank-0-only or time-based loop would deadlock); count derived from the max-over-ranks step time
    tcont = torch.tensor([total_ms / args.steps], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tcont, op=dist.ReduceOp.MAX)
    ncont = int(min(20000, max(1, 1200.0 / max(float(tcont[0]), 1e-3))))
    for _ in range(ncont):
        for _, k in kernels:
            k()
    barrier()
    clk = clocks.stop() if clocks else None

    # end to end through the host-pointer API (copies inside the timed region): pinned buffers, then pageable ones
    def time_host(step):
        if step is None:
            return 0.0
        step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        barrier()
        return (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = time_host(e2e_step)
    e2e_pg_ms = time_host(e2e_pageable_step)

    t = torch.tensor([total_ms, e2e_ms, e2e_pg_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_max, e2e_pg_max = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = total_ms / args.steps
    strong = w["kind"] == "slab"     # one array split over the ranks: total work is fixed
    value = flops_per_step(w) * (1 if strong else world) / (ms_per_step * 1e-3) / 1e9

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        jdom = int(np.argmax(per_kernel_ms))
        if w["kind"] == "slab":
            abytes = algorithmic_bytes(w) * 3 // world         # three local passes over this rank's share
            dom_ms = per_kernel_ms[0]
            dom_label = "3 local passes + exchange (whole step)"
        elif w["kind"] == "nd":
            abytes = algorithmic_bytes(w) * len(w["dims"])     # the fftnd call = ndims axis-pass launches
            dom_ms = per_kernel_ms[0]
            dom_label = "kf axis pass x%d" % len(w["dims"])
        else:
            abytes = algorithmic_bytes(w)
            dom_ms = per_kernel_ms[jdom]
            dom_label = kernels[jdom][0]
        achieved = abytes / (dom_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": DTYPE_NAME[w["tname"]], "data": "synthetic",
            "config": config_dict(w, args.workload, world),
            "kernel_ms": {k[0]: ms for k, ms in zip(kernels, per_kernel_ms)},
            "roofline": {"bound": "hbm", "kernel": dom_label, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                         "algorithmic_bytes": abytes, "traffic": known_traffic(args.workload),
                         "traffic_source": "static: dram bytes of the committed ncu capture (profiles/traffic.json), not measured in this run",
                         "all_kernels": {k[0]: {"ms": ms, "frac": algorithmic_bytes(w) / (ms * 1e-3) / 1e9 / peak}
                                         for k, ms in zip(kernels, per_kernel_ms)} if w["kind"] in ("real", "c2c") else None},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if e2e_step is not None:
            api = "kiss_fftr_batch+kiss_fftri_batch" if w["kind"] == "real" else "kiss_fft_batch"
            line["e2e"] = {"value": flops_per_step(w) * world / (e2e_max * 1e-3) / 1e9, "unit": "GFLOP/s",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_max,
                           "api": api, "host_buffers": "pinned (cudaHostAlloc)"}
            line["e2e_pageable"] = {"value": flops_per_step(w) * world / (e2e_pg_max * 1e-3) / 1e9, "unit": "GFLOP/s",
                                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_pg_max,
                                    "api": api, "host_buffers": "pageable (malloc / numpy), staged by the library"}
        else:
            line["e2e"] = None

    # ---- the other BASELINE configs, measured in the same run (never allowed to take the headline down) ----
    extras = args.workload == "r2c4096" and not args.no_configs
    # watchdog: whatever happens below (a peer that never signals, a dead CUDA context, a collective that hangs), the
    # headline line measured above is printed and every rank exits 0 after BENCH_EXTRAS_TIMEOUT seconds
    state = {"emitted": False}

    def bail():
        if rank == 0 and line is not None and not state["emitted"]:
            state["emitted"] = True
            line.setdefault("configs", {})["watchdog"] = "the extra configs did not finish within the time limit; headline numbers are complete"
            emit(line)
        os._exit(0)
    import threading
    dog = threading.Timer(float(os.environ.get("BENCH_EXTRAS_TIMEOUT", "600")), bail)
    dog.daemon = True
    dog.start()
    if extras:
        # release the headline's buffers first: fftnd 1024^3 needs 16 GiB, the slabs up to 16 GiB per rank
        kernels = e2e_step = e2e_pageable_step = None
        d_x = d_X = d_y = h_x = h_X = h_y = p_x = p_X = p_y = None
        torch.cuda.empty_cache()
    try:
        _run_extras(args, extras, line, rank, world, dist, barrier, w)
    except BaseException as exc:          # incl. a CUDA context lost to a trapped kernel
        if line is not None:
            line.setdefault("configs", {})["error"] = "%s: %s" % (type(exc).__name__, exc)
    if rank == 0 and not state["emitted"]:
        state["emitted"] = True
        emit(line)
    dog.cancel()
    try:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
    except BaseException:
        pass


def _run_extras(args, extras, line, rank, world, dist, barrier, w):
    import torch
    if extras:
        configs = {}
        if world == 1:
            for nm in ("c2c1024", "c2c1000", "c2c1155", "z2z1000", "z2z1155", "q15_2048", "q31_2048", "fftnd1024"):
                try:
                    configs[nm] = measure_config(nm, args, barrier)
                except Exception as exc:
                    configs[nm] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        else:
            t1 = None
            try:                                   # the 1-GPU time of the same transform, rank 0 alone
                if rank == 0:
                    t1 = measure_config("fftnd1024", args, lambda: torch.cuda.synchronize())
                    configs["fftnd1024_rank0_alone"] = t1
                dist.barrier()
            except Exception as exc:
                configs["fftnd1024_rank0_alone"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            tt = torch.tensor([t1["ms"] if (t1 and "ms" in t1) else 0.0], device="cuda", dtype=torch.float64)
            dist.broadcast(tt, 0)
            for nm in ("slab1024", "slab1024nccl"):
                try:
                    configs[nm] = measure_slab(nm, args, dist, rank, world, float(tt[0]) or None)
                except Exception as exc:
                    configs[nm] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        if line is not None:
            line["configs"] = configs

    if rank == 0 and world == 1:
        try:
            line["cpu_baseline"] = best_cpu_reference(w, 3, detail=True)
        except Exception as exc:   # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"error": str(exc)}


def best_cpu_reference(w, k, detail=False):
    """best of k runs of the CPU reference arm (the box's other tenants make single runs vary by 20-30 %)"""
    best = None
    for i in range(k):
        gf, info = cpu_reference(w, reps=1, detail=detail and i == 0)
        if best is None or gf > best["value"]:
            keep = best.get("other_modes_gflops") if best else None
            best = dict(info)
            if keep and "other_modes_gflops" not in best:
                best["other_modes_gflops"] = keep
        elif "other_modes_gflops" in info and "other_modes_gflops" not in best:
            best["other_modes_gflops"] = info["other_modes_gflops"]
    best["best_of"] = k
    return best


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    for _ in range(max(1, args.warmup)):
        cpu_reference(w, reps=1)
    vals, info = [], None
    for _ in range(args.steps):
        gf, i = cpu_reference(w, reps=1)
        vals.append(gf)
        if info is None or gf >= max(vals):
            info = i
    value = float(max(vals))                        # best of K steps, like the repo arm's own cpu_baseline
    info = dict(info, value=value, best_of=args.steps, median=float(np.median(vals)))
    dt = flops_per_step(w) / (value * 1e9)          # one full step of the workload at the measured rate
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NAME[w["tname"]], "data": "synthetic",
            "config": config_dict(w, args.workload, world),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: anything libraries print there (NCCL banners, warnings) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="r2c4096", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-configs", action="store_true", help="headline workload only (skip the `configs` object)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
