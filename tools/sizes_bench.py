#!/usr/bin/env python
"""times kiss_fft_batch_dev over a list of lengths (development aid): which kernel serves a length and how close it runs
to the HBM roofline.  usage: python tools/sizes_bench.py [tname] [n1 n2 ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kissfft_b200  # noqa: E402

tname = sys.argv[1] if len(sys.argv) > 1 else "float"
sizes = [int(a) for a in sys.argv[2:]] or [120, 240, 512, 1800, 3000, 4096, 8192, 16384, 65536]
lib = kissfft_b200.get(tname)
tdt = {"float": torch.float32, "double": torch.float64, "int16_t": torch.int16, "int32_t": torch.int32}[tname]
for n in sizes:
    batch = max(1, (256 << 20) // (n * 2 * torch.empty((), dtype=tdt).element_size()))
    x = torch.zeros((batch, n, 2), dtype=tdt, device="cuda")
    y = torch.empty_like(x)
    cfg = lib.alloc(n)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.fft_batch_dev(cfg, x, y, batch, n, n, 1, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.fft_batch_dev(cfg, x, y, batch, n, n, 1, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = 2 * x.numel() * x.element_size() / ms * 1e-6
    print(json.dumps({"type": tname, "nfft": n, "batch": batch, "ms": round(ms, 4), "GBps": round(gbs, 1), "plan_kind": lib.plan_kind(n)}))
    lib.free(cfg)
    del x, y
