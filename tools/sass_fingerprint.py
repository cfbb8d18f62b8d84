#!/usr/bin/env python
"""md5 of the SASS of every product library (cuobjdump -sass, address-only lines dropped).

A change that is meant to leave the shipped kernels untouched (a new plan option that defaults to off, a refactoring of
the templates) can be checked without a GPU: the fingerprints must not move.

    python tools/sass_fingerprint.py                 # print
    python tools/sass_fingerprint.py --check FILE    # compare with a saved listing, exit 1 on a difference
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TYPES = ("float", "double", "int16_t", "int32_t")


def fingerprint(tname):
    lib = os.path.join(ROOT, "kissfft_b200", "lib", "libkissfft-%s.so" % tname)
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    h = hashlib.md5()
    for line in out.splitlines():
        if re.match(r"^\s*/\*[0-9a-f]*\*/\s*$", line):
            continue
        h.update(line.encode() + b"\n")
    return h.hexdigest()


def main():
    now = {t: fingerprint(t) for t in TYPES}
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        saved = dict(line.split() for line in open(sys.argv[2]) if line.strip() and not line.startswith("#"))
        bad = [t for t in TYPES if saved.get(t) != now[t]]
        for t in TYPES:
            print("%-8s %s %s" % (t, now[t], "CHANGED" if t in bad else "same"))
        sys.exit(1 if bad else 0)
    for t in TYPES:
        print("%-8s %s" % (t, now[t]))


if __name__ == "__main__":
    main()
