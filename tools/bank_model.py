#!/usr/bin/env python
"""Shared-memory wavefront model of a fused plan's exchange traffic (development aid, CPU only).

Re-states PlanDesc's index math (kissfft_b200/csrc/kf_plan.h) and counts, per transform, the shared-memory wavefronts of
every warp-wide LDS/STS of the exchange buffers for 8-byte elements: a warp access is served as two half-warps, each
costing max(multiplicity of distinct 8-byte slots per bank pair) wavefronts.  Compared with ncu
(l1tex__data_pipe_lsu_wavefronts_mem_shared_op_{ld,st}) it reproduces the measured conflict counts, so padding schemes
can be screened without a GPU.

    python tools/bank_model.py            # the round-1 R2C / C2R 4096-point plans
"""
import sys
from functools import reduce


class Plan:
    def __init__(self, N, p, glen, team, logpad, skew=None):
        self.N, self.p, self.glen, self.team, self.logpad = N, list(p), list(glen), team, logpad
        self.L, self.G = len(p), len(glen)
        self.skew = skew

    def F(self, s):
        return reduce(lambda a, b: a * b, self.p[:s], 1)

    def m(self, s):
        return self.N // (self.F(s) * self.p[s])

    def s_hi(self, g):
        return self.L - 1 - sum(self.glen[:g])

    def s_lo(self, g):
        return self.s_hi(g) - self.glen[g] + 1

    def R(self, g):
        return reduce(lambda a, b: a * b, self.p[self.s_lo(g):self.s_hi(g) + 1], 1)

    def Flo(self, g):
        return self.F(self.s_lo(g))

    def items(self, g):
        return self.N // self.R(g)

    def W(self, g, s):
        return reduce(lambda a, b: a * b, self.p[self.s_lo(g):s], 1)

    def digit(self, g, s, e):
        return (e // self.W(g, s)) % self.p[s]

    def kout(self, g, e):
        return sum(self.digit(g, j, e) * self.m(j) for j in range(self.s_lo(g), self.s_hi(g) + 1))

    def phys(self, a):
        if self.skew is not None:
            return self.skew(a)
        return a if self.logpad >= 31 else a + (a >> self.logpad)

    # addresses (element units) of work item w of group g
    def rd(self, g, w):
        Flo, R = self.Flo(g), self.R(g)
        off, kp = w % Flo, w // Flo
        return [self.phys(kp * Flo * R + off + e * Flo) for e in range(R)]

    def wr(self, g, w):
        Flo, R = self.Flo(g), self.R(g)
        off, kp = w % Flo, w // Flo
        return [self.phys((kp + self.kout(g, e)) * Flo + off) for e in range(R)]


ESIZE = 8     # bytes per complex element: 8 float, 16 double, 4 Q15, 8 Q31


def wavefronts(addrs):
    """addrs: 32 element indices (None = inactive lane) of one warp-wide access of ESIZE-byte elements.  The access is
    served in phases of 128 / ESIZE lanes; a phase costs as many wavefronts as distinct elements share one bank group."""
    lanes = 128 // ESIZE
    total = 0
    for i in range(0, 32, lanes):
        banks = {}
        for a in addrs[i:i + lanes]:
            if a is not None:
                banks.setdefault(a % lanes, set()).add(a)
        total += max((len(v) for v in banks.values()), default=0)
    return total


def count(plan, g, kind, item_of_thread):
    """wavefronts per transform of group g's exchange loads ('rd') or stores ('wr'); item_of_thread(t, it) -> item or None"""
    R, team = plan.R(g), plan.team
    its = 0
    total = 0
    while True:
        any_on = False
        for w0 in range(0, team, 32):
            lanes = [item_of_thread(t, its) for t in range(w0, w0 + 32)]
            if all(l is None for l in lanes):
                continue
            any_on = True
            per_lane = [(plan.rd(g, l) if kind == "rd" else plan.wr(g, l)) if l is not None else None for l in lanes]
            for e in range(R):
                total += wavefronts([pl[e] if pl is not None else None for pl in per_lane])
        if not any_on:
            break
        its += 1
    return total


def pitch(plan):
    n = plan.phys(plan.N - 1) + 1
    return n + 1 if n % 2 == 0 else n


def count_cta(plan, g, kind, tpc):
    """like count(), but over a whole CTA of tpc teams (tid = team * plan.team + t): warps may straddle two transforms when
    the team size is not a multiple of 32.  Returns wavefronts per transform (a float)."""
    R, T, n = plan.R(g), plan.team, plan.items(g)
    P = pitch(plan)
    nthreads = T * tpc
    iters = -(-n // T)
    total = 0
    for it in range(iters):
        for w0 in range(0, nthreads, 32):
            lanes = []
            for tid in range(w0, min(w0 + 32, nthreads)):
                team, t = divmod(tid, T)
                w = t + it * T
                lanes.append((team, w) if w < n else None)
            lanes += [None] * (32 - len(lanes))
            if all(l is None for l in lanes):
                continue
            per = [None if l is None else [l[0] * P + a for a in (plan.rd(g, l[1]) if kind == "rd" else plan.wr(g, l[1]))] for l in lanes]
            for e in range(R):
                total += wavefronts([pl[e] if pl is not None else None for pl in per])
    return total / tpc


def report_cta(name, plan, tpc):
    ideal = plan.N * ESIZE / 128
    ld = sum(count_cta(plan, g, "rd", tpc) for g in range(1, plan.G))
    st = sum(count_cta(plan, g, "wr", tpc) for g in range(plan.G - 1))
    print("%-34s radices=%s groups=%s team=%d tpc=%d logpad=%d: loads %.0f stores %.0f (ideal %.0f each per exchange, %d exchanges)" % (
        name, plan.p, plan.glen, plan.team, tpc, plan.logpad, ld, st, ideal, plan.G - 1))
    return ld + st


def plain(plan, g):
    n = plan.items(g)
    return lambda t, it: (t + it * plan.team) if (t + it * plan.team) < n else None


def paired(plan, g, second):
    """items (u, m-u); (0, m/2) for u == 0.  second=False -> item a of the pair, True -> item b"""
    m = plan.items(g)
    half = m // 2

    def f(t, it):
        u = t + it * plan.team
        if u >= half:
            return None
        if not second:
            return u
        return half if u == 0 else m - u
    return f


def report(name, plan, mode):
    G = plan.G
    ideal = -(-plan.N * ESIZE // 128)
    print("== %s  N=%d radices=%s groups=%s team=%d logpad=%d" % (name, plan.N, plan.p, plan.glen, plan.team, plan.logpad))
    tot_ld = tot_st = 0
    for g in range(G):
        if g > 0:      # exchange loads
            if mode == "r2c" and g == G - 1:
                n = count(plan, g, "rd", paired(plan, g, False)) + count(plan, g, "rd", paired(plan, g, True))
            else:
                n = count(plan, g, "rd", plain(plan, g))
            print("   group %d loads  : %4d wavefronts (ideal %d)" % (g, n, ideal))
            tot_ld += n
        if g < G - 1:  # exchange stores
            if mode == "c2r" and g == 0:
                n = count(plan, g, "wr", paired(plan, g, False)) + count(plan, g, "wr", paired(plan, g, True))
            else:
                n = count(plan, g, "wr", plain(plan, g))
            print("   group %d stores : %4d wavefronts (ideal %d)" % (g, n, ideal))
            tot_st += n
    print("   exchange loads %d, stores %d per transform (input-stage reads not included)" % (tot_ld, tot_st))
    return tot_ld, tot_st


def set_esize(n):
    global ESIZE
    ESIZE = n


if __name__ == "__main__":
    report("R2C 4096 (round-1 plan)", Plan(2048, [4, 2, 4, 4, 4, 4], [2, 2, 2], 128, 4), "r2c")
    report("C2R 4096 (round-1 plan)", Plan(2048, [4, 4, 4, 4, 4, 2], [2, 2, 2], 128, 4), "c2r")
    report("C2C f32 1024", Plan(1024, [2, 4, 4, 4, 4, 2], [3, 3], 32, 5), "c2c")
    report("C2C f32 2048", Plan(2048, [4, 2, 4, 4, 4, 4], [2, 2, 2], 128, 4), "c2c")
    report("C2C f32 1000", Plan(1000, [5, 5, 5, 4, 2], [3, 2], 40, 5), "c2c")
    report("C2C f32 1155", Plan(1155, [3, 11, 5, 7], [2, 2], 35, 5), "c2c")
    set_esize(16)
    report("C2C f64 1000", Plan(1000, [4, 2, 5, 5, 5], [1, 1, 1, 2], 200, 3), "c2c")
    report("C2C f64 1155", Plan(1155, [3, 5, 7, 11], [1, 1, 2], 105, 5), "c2c")
    set_esize(4)
    report("C2C Q15 2048", Plan(2048, [4, 4, 4, 4, 4, 2], [2, 2, 2], 128, 4), "c2c")
    print("-- whole-CTA counts (team sizes that are not a multiple of 32)")
    set_esize(8)
    report_cta("C2C f32 1000", Plan(1000, [5, 5, 5, 4, 2], [3, 2], 40, 5), 3)
    report_cta("C2C f32 1155", Plan(1155, [3, 11, 5, 7], [2, 2], 35, 5), 3)
    set_esize(16)
    report_cta("C2C f64 1000", Plan(1000, [4, 2, 5, 5, 5], [1, 1, 1, 2], 200, 3), 1)
    report_cta("C2C f64 1155", Plan(1155, [3, 5, 7, 11], [1, 1, 2], 105, 5), 1)
