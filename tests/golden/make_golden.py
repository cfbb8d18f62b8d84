"""Generates tests/golden/golden_<type>.npz from the UNMODIFIED reference compiled by `make -C oracle ref`
(oracle/_ref/libkissfft-<type>.so, built from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each file stores, per case, the exact input and the reference's exact output, so the fixtures pin the oracle
(tests/test_oracle_pin.py) and the CUDA path (tests/test_gpu_golden.py) on machines where /root/reference and
oracle/_ref do not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.loader import TYPES, Reference, build_reference, random_input  # noqa: E402

C2C = [(16, 0), (30, 0), (120, 1), (1000, 0), (1024, 0), (1155, 0), (1155, 1), (2048, 0), (2048, 1), (74, 0)]
REAL = [30, 120, 1000, 4096]
ND = [((4, 3), 0), ((2, 3, 4), 0), ((8, 8, 8), 1), ((30, 20, 12), 0)]
NDR = [(6, 8), (4, 6, 10)]


def main():
    assert build_reference(), "needs /root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    for t in TYPES:
        r = Reference(t)
        d = {}
        for n, inv in C2C:
            x = random_input(t, (2, n), 9000 + n + inv)
            d["c2c_%d_%d_in" % (n, inv)] = x
            d["c2c_%d_%d_out" % (n, inv)] = r.fft(x, inv)
        for n in REAL:
            x = random_input(t, (2, n), 9100 + n, complex_=False)
            X = r.fftr(x)
            d["r2c_%d_in" % n] = x
            d["r2c_%d_out" % n] = X
            S = X if t in ("float", "double") else random_input(t, (2, n // 2 + 1), 9200 + n)
            d["c2r_%d_in" % n] = S
            d["c2r_%d_out" % n] = r.fftri(S)
        for dims, inv in ND:
            x = random_input(t, dims, 9300 + len(dims))
            key = "x".join(map(str, dims))
            d["nd_%s_%d_in" % (key, inv)] = x
            d["nd_%s_%d_out" % (key, inv)] = r.fftnd(x, inv)
        for dims in NDR:
            x = random_input(t, dims, 9400 + len(dims), complex_=False)
            key = "x".join(map(str, dims))
            X = r.fftndr(x)
            d["ndr_%s_in" % key] = x
            d["ndr_%s_out" % key] = X
            S = X if t in ("float", "double") else random_input(t, X.shape[:-1], 9500 + len(dims))
            d["ndri_%s_in" % key] = S
            d["ndri_%s_out" % key] = r.fftndri(S)
        path = os.path.join(here, "golden_%s.npz" % t)
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), "bytes", len(d) // 2, "cases")


if __name__ == "__main__":
    main()
