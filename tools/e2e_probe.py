#!/usr/bin/env python
"""sweeps the host-pointer pipeline knobs (KISSFFT_HOST_LANES x KISSFFT_CHUNK_MIB) for pinned and pageable caller buffers
on the headline workload (R2C + C2R f32 4096 x 32768); development aid.  usage: python tools/e2e_probe.py [lanes...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kissfft_b200  # noqa: E402

lib = kissfft_b200.get("float")
n, b = 4096, 32768
nb = n // 2 + 1
cf, ci = lib.allocr(n, False), lib.allocr(n, True)
h_x = (torch.rand((b, n)) * 2 - 1).pin_memory()
h_X = torch.empty((b, nb, 2)).pin_memory()
h_y = torch.empty((b, n)).pin_memory()
p_x = np.array(h_x.numpy(), copy=True)
p_X = np.empty((b, nb, 2), np.float32)
p_y = np.empty((b, n), np.float32)
flops = 2 * 2.5 * n * np.log2(n) * b


def run(x, X, y, reps=4):
    lib.fftr_batch(cf, x, X, b)
    lib.fftri_batch(ci, X, y, b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.fftr_batch(cf, x, X, b)
        lib.fftri_batch(ci, X, y, b)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


lanes = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12]
for chunk in (2, 4, 8, 16, 32):
    for ln in lanes:
        os.environ["KISSFFT_HOST_LANES"] = str(ln)
        os.environ["KISSFFT_CHUNK_MIB"] = str(chunk)
        ms_pin = run(h_x, h_X, h_y)
        ms_pg = run(p_x, p_X, p_y)
        print(json.dumps({"lanes": ln, "chunk_mib": chunk, "pinned_ms": round(ms_pin, 2), "pageable_ms": round(ms_pg, 2),
                          "pinned_gflops": round(flops / ms_pin / 1e6, 1), "pageable_gflops": round(flops / ms_pg / 1e6, 1)}), flush=True)
assert np.allclose(p_y[:64] / n, p_x[:64], atol=1e-4)
