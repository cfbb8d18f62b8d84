#!/usr/bin/env python
"""Host-pointer pipeline probe: times kiss_fftr_batch + kiss_fftri_batch (pinned host buffers, H2D + kernel + D2H inside
the call) for several KISSFFT_CHUNK_MIB settings.  Development aid; run on the GPU box."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kissfft_b200  # noqa: E402


def main():
    lib = kissfft_b200.get("float")
    nfft, batch = 4096, 32768
    x = torch.rand((batch, nfft), dtype=torch.float32).mul_(2).sub_(1).pin_memory()
    X = torch.empty((batch, nfft // 2 + 1, 2), dtype=torch.float32).pin_memory()
    y = torch.empty_like(x).pin_memory()
    cf, ci = lib.allocr(nfft, False), lib.allocr(nfft, True)
    for mib in [int(a) for a in sys.argv[1:]] or [32, 16, 8, 4]:
        os.environ["KISSFFT_CHUNK_MIB"] = str(mib)
        ts = []
        for it in range(6):
            t0 = time.perf_counter()
            lib.fftr_batch(cf, x.data_ptr(), X.data_ptr(), batch)
            lib.fftri_batch(ci, X.data_ptr(), y.data_ptr(), batch)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[1:])
        gb = (x.numel() * 4 + X.numel() * 4) * 2 / 1e9
        print(json.dumps({"chunk_mib": mib, "ms_med": ts[len(ts) // 2] * 1e3, "ms_best": ts[0] * 1e3,
                          "GBps_each_way": gb / 2 / ts[len(ts) // 2]}))


if __name__ == "__main__":
    main()
