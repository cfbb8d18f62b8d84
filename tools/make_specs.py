#!/usr/bin/env python
"""Generates tuner specs (see tune_gen.py) for the BASELINE configurations: candidate radix orders, register
groupings, team / transforms-per-CTA / occupancy / input-ring depth.  Development aid."""
import itertools
import json
import os
import random
import sys

OUT = "/tmp/specs"


def kiss_factors(n):
    p, out, root = 4, [], int(n ** 0.5)
    while n > 1:
        while n % p:
            p = 2 if p == 4 else (3 if p == 2 else p + 2)
            if p > root:
                p = n
        n //= p
        out.append(p)
    return out


def groupings(rad, rmax, gmax):
    """contiguous splits of rad (outermost..innermost) into groups, returned FIRST-EXECUTED (innermost) first"""
    L = len(rad)
    res = []
    for G in range(1, gmax + 1):
        for cuts in itertools.combinations(range(1, L), G - 1):
            b = [0] + list(cuts) + [L]
            parts = [rad[b[i]:b[i + 1]] for i in range(G)]
            R = [eval("*".join(map(str, p))) for p in parts]
            if max(R) <= rmax:
                res.append(([len(p) for p in parts][::-1], R[::-1]))
    return res


def orders(rad, fixed, limit=6):
    if fixed:
        return [rad]
    perms = sorted(set(itertools.permutations(rad)))
    random.Random(1).shuffle(perms)
    keep = [tuple(rad)] + [p for p in perms if p != tuple(rad)][:limit - 1]
    return [list(p) for p in keep]


def spec(name, tname, mode, nfft, fixed, rmax, alt_radices=(), ncand=14, tpcs=(1, 2, 4), stages=(0, 2)):
    base = kiss_factors(nfft)
    cands = []
    for rad0 in [base] + [list(a) for a in alt_radices]:
        for rad in orders(rad0, fixed and rad0 == base):
            gs = groupings(rad, rmax, 4)
            if not gs:
                continue
            gmin = min(len(g[0]) for g in gs)
            for glen, R in gs:
                if len(glen) > gmin:
                    continue
                for team in sorted({nfft // max(R), nfft // min(R)}):
                    if team > 512 or team < 16:
                        continue
                    cands.append((rad, glen, R, team))
    random.Random(2).shuffle(cands)
    cands.sort(key=lambda c: len(c[1]))
    variants = []
    seen = set()
    for rad, glen, R, team in cands[:ncand]:
        for tpc in tpcs:
            thr = team * tpc
            if thr < 32 or thr > 512:
                continue
            for mb in sorted({max(1, 256 // thr), max(1, 512 // thr), max(1, 768 // thr)}):
                for ns in stages:
                    key = (tuple(rad), tuple(glen), team, tpc, mb, ns)
                    if key in seen:
                        continue
                    seen.add(key)
                    variants.append(dict(radices=rad, groups=glen, team=team, tpc=tpc, logpad=4, minblocks=mb, nstage=ns))
    return dict(name=name, type=tname, mode=mode, nfft=nfft, radices=base, variants=variants)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "round1"
    if which == "round1":
        S = [
            spec("f32_r2c4096", "float", "R2C", 2048, False, 32, alt_radices=([4, 4, 4, 4, 2, 2, 2], [4, 4, 4, 2, 2, 2, 2, 2])),
            spec("f32_c2r4096", "float", "C2R", 2048, False, 32, alt_radices=([4, 4, 4, 4, 2, 2, 2],)),
            spec("f32_1000", "float", "C2C", 1000, False, 32),
            spec("f32_1155", "float", "C2C", 1155, False, 35),
            spec("f64_1000", "double", "C2C", 1000, False, 16),
            spec("f64_1155", "double", "C2C", 1155, False, 16),
            spec("q15_2048", "int16_t", "C2C", 2048, True, 32),
            spec("q31_2048", "int32_t", "C2C", 2048, True, 32),
        ]
    else:   # "bigR": fewer shared-memory exchanges, more registers per thread
        S = [
            spec("f32_r2c4096_R", "float", "R2C", 2048, False, 64, ncand=8, tpcs=(1, 2), stages=(0, 2)),
            spec("f32_c2r4096_R", "float", "C2R", 2048, False, 64, ncand=8, tpcs=(2,), stages=(0, 2)),
            spec("f32_1000_R", "float", "C2C", 1000, False, 40, ncand=8, tpcs=(1, 2), stages=(0, 2)),
            spec("f32_1155_R", "float", "C2C", 1155, False, 35, alt_radices=([5, 7, 3, 11], [3, 11, 5, 7], [7, 5, 11, 3]), ncand=10, tpcs=(1, 2), stages=(0, 2)),
            spec("f64_1000_R", "double", "C2C", 1000, False, 25, ncand=8, tpcs=(1, 2), stages=(0, 2)),
            spec("f64_1155_R", "double", "C2C", 1155, False, 21, alt_radices=([5, 11, 3, 7], [3, 7, 5, 11]), ncand=10, tpcs=(1, 2), stages=(0, 2)),
            spec("q15_2048_R", "int16_t", "C2C", 2048, True, 64, ncand=8, tpcs=(1, 2), stages=(0, 2)),
            spec("q31_2048_R", "int32_t", "C2C", 2048, True, 64, ncand=8, tpcs=(1, 2), stages=(0, 2)),
        ]
    for s in S:
        json.dump(s, open(os.path.join(OUT, s["name"] + ".json"), "w"))
        print(s["name"], len(s["variants"]), sorted({(tuple(v["radices"]), tuple(v["groups"]), v["team"]) for v in s["variants"]})[:3])
