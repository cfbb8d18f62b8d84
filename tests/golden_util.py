"""Access to the committed golden vectors (generated from the compiled reference by tests/golden/make_golden.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load(tname):
    return np.load(os.path.join(HERE, "golden", "golden_%s.npz" % tname))


def cases(tname, prefix):
    """yields (key_suffix, input, output) for every case whose key starts with prefix + '_'"""
    g = load(tname)
    for k in g.files:
        if k.startswith(prefix + "_") and k.endswith("_in"):
            base = k[: -len("_in")]
            yield base[len(prefix) + 1:], g[k], g[base + "_out"]
