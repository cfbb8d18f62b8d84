#!/usr/bin/env python
"""fused fast convolution vs the same pipeline assembled from three launches (development aid / evidence for DESIGN.md)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kissfft_b200  # noqa: E402

lib = kissfft_b200.get("float")
for nfft, nimp in ((1024, 129), (2048, 257), (4096, 513)):
    rng = np.random.default_rng(1)
    imp = (rng.uniform(-1, 1, size=(nimp, 2)) / nimp).astype(np.float32)
    cfg, n, ngood = lib.fastconv_alloc(imp, nfft)
    nblocks = (512 << 20) // (nfft * 8)
    nsamp = (nblocks - 1) * ngood + nfft
    x = torch.rand((nsamp, 2), device="cuda") * 2 - 1
    y = torch.zeros_like(x)
    st = torch.cuda.current_stream().cuda_stream

    def timeit(fn, it=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it

    ms_fused = timeit(lambda: lib.fastconv_dev(cfg, x, y, nsamp, st))
    # unfused: forward FFT of the overlapping blocks, pointwise multiply, inverse FFT, keep ngood samples
    cf, ci = lib.alloc(nfft, False), lib.alloc(nfft, True)
    X = torch.empty((nblocks, nfft, 2), device="cuda")
    Y = torch.empty_like(X)
    H = torch.rand((nfft, 2), device="cuda")
    y2 = torch.zeros((nblocks, ngood, 2), device="cuda")

    def unfused():
        lib.fft_batch_dev(cf, x, X, nblocks, ngood, nfft, 1, st)
        Xc, Hc = torch.view_as_complex(X), torch.view_as_complex(H)
        torch.view_as_complex(Y).copy_(Xc * Hc)
        lib.fft_batch_dev(ci, Y, Y, nblocks, nfft, nfft, 1, st)
        y2.copy_(Y[:, :ngood])

    ms_unfused = timeit(unfused)
    abytes = nblocks * (nfft + ngood) * 8
    print(json.dumps({"nfft": nfft, "nimp": nimp, "blocks": nblocks, "fused_ms": round(ms_fused, 4), "unfused_ms": round(ms_unfused, 4),
                      "speedup": round(ms_unfused / ms_fused, 2), "fused_GBps_algorithmic": round(abytes / ms_fused * 1e-6, 1)}))
    lib.fastconv_free(cfg)
