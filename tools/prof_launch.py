#!/usr/bin/env python
"""Launches the dominant kernel of every BASELINE config a few times (profiling aid, meant to run under ncu):

    ncu --set full --clock-control none --import-source on -k regex:kf_ -o gpurun_out/all python tools/prof_launch.py [names...]

names: c2c1024 r2c4096 c2r4096 c2c1000 c2c1155 z2z1000 z2z1155 q15_2048 q31_2048 fftnd1024 (default: all but fftnd1024).
Each kernel is launched `reps` times (env PROF_REPS, default 2) on random data of the BASELINE shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kissfft_b200  # noqa: E402

TDT = {"float": torch.float32, "double": torch.float64, "int16_t": torch.int16, "int32_t": torch.int32}
C2C = {"c2c1024": ("float", 1024, 65536), "c2c1000": ("float", 1000, 100000), "c2c1155": ("float", 1155, 100000),
       "z2z1000": ("double", 1000, 100000), "z2z1155": ("double", 1155, 100000), "q15_2048": ("int16_t", 2048, 65536),
       "q31_2048": ("int32_t", 2048, 65536)}
ALL = ["c2c1024", "r2c4096", "c2r4096", "c2c1000", "c2c1155", "z2z1000", "z2z1155", "q15_2048", "q31_2048"]


def rnd(shape, tname):
    if tname in ("float", "double"):
        return (torch.rand(shape, dtype=TDT[tname], device="cuda") * 2 - 1)
    half = 16383 if tname == "int16_t" else 1073741823
    return torch.randint(-half, half + 1, shape, dtype=TDT[tname], device="cuda")


def main():
    names = sys.argv[1:] or ALL
    reps = int(os.environ.get("PROF_REPS", "2"))
    st = torch.cuda.current_stream().cuda_stream
    for name in names:
        if name in C2C:
            tname, n, batch = C2C[name]
            lib = kissfft_b200.get(tname)
            x = rnd((batch, n, 2), tname)
            y = torch.empty_like(x)
            cfg = lib.alloc(n)
            for _ in range(reps):
                lib.fft_batch_dev(cfg, x, y, batch, n, n, 1, st)
        elif name in ("r2c4096", "c2r4096"):
            lib = kissfft_b200.get("float")
            n, batch = 4096, 32768
            t = rnd((batch, n), "float")
            f = rnd((batch, n // 2 + 1, 2), "float")
            cfg = lib.allocr(n, name == "c2r4096")
            for _ in range(reps):
                if name == "r2c4096":
                    lib.fftr_batch_dev(cfg, t, f, batch, n, n // 2 + 1, st)
                else:
                    lib.fftri_batch_dev(cfg, f, t, batch, n // 2 + 1, n, st)
        elif name.startswith("fftnd"):
            d = int(name[5:])
            lib = kissfft_b200.get("float")
            x = rnd((d, d, d, 2), "float")
            y = torch.empty_like(x)
            cfg = lib.allocnd((d, d, d))
            for _ in range(reps):
                lib.fftnd_dev(cfg, x, y, None, st)
        else:
            raise SystemExit("unknown workload " + name)
        torch.cuda.synchronize()
        lib.free(cfg)
        print("launched", name, flush=True)


if __name__ == "__main__":
    main()
