// kf_twtab.h -- host-side construction of the per-group stage-twiddle tables of a fused plan (see
// PlanDesc::slot in kf_plan.h).  Every entry is an exact copy of an entry of the plan's twiddle array
// (kiss_fft.c:361-367), so the fixed-point builds stay bit-identical; only the order in memory changes, from
// "indexed by q*k*fstride" (kiss_fft.c:54-66) to "indexed by (butterfly slot, work item)".
#pragma once
#include <vector>

#include "kf_body.h"

namespace kf {

// groups 1..G-1: table[gtw_offset(g) + slot*items(g) + w] = tw[q * F_s * (k'(w) + kabove(g, s, e))]
template <class A, class PT>
std::vector<typename A::C> build_gtw(const typename A::C* h_tw)
{
    constexpr PlanDesc D = PT::D;
    std::vector<typename A::C> tab((size_t)(D.gtw_total() > 0 ? D.gtw_total() : 1));
    for (int g = 1; g < D.G; ++g)
        for (int s = D.s_hi(g); s >= D.s_lo(g); --s)
            for (int e = 0; e < D.R(g); ++e) {
                if (D.digit(g, s, e) != 0) continue;
                for (int q = 1; q < D.p[s]; ++q) {
                    const int slot = D.slot(g, s, e) + q - 1;
                    for (int w = 0; w < D.items(g); ++w) {
                        const int kp = w / D.Flo(g);
                        const long long idx = (long long)q * D.F(s) * (kp + D.kabove(g, s, e));
                        tab[(size_t)D.gtw_offset(g) + (size_t)slot * D.items(g) + w] = h_tw[idx];
                    }
                }
            }
    return tab;
}

// group 0 (k' == 0): P.g0tw[slot] = tw[q * F_s * kabove(0, s, e)]
template <class A, class PT>
void fill_g0tw(KParams<A>& P, const typename A::C* h_tw)
{
    constexpr PlanDesc D = PT::D;
    for (int s = D.s_hi(0); s >= D.s_lo(0); --s)
        for (int e = 0; e < D.R(0); ++e) {
            if (D.digit(0, s, e) != 0) continue;
            for (int q = 1; q < D.p[s]; ++q)
                P.g0tw[D.slot(0, s, e) + q - 1] = A::load(h_tw[(long long)q * D.F(s) * D.kabove(0, s, e)]);
        }
    // split-twiddle constants tw[q * F_s * kabove] for every non-zero upper-digit combination
    if (D.twmode == 1)
        for (int g = 1; g < D.G; ++g)
            for (int s = D.s_hi(g); s >= D.s_lo(g); --s)
                for (int up = 1; up < D.nupper(g, s); ++up)
                    for (int q = 1; q < D.p[s]; ++q)
                        P.ctw[D.cslot(g, s, up, q)] =
                            A::load(h_tw[(long long)q * D.F(s) * D.kabove(g, s, D.upper_base(g, s, up))]);
}

// one direction of the fused fast convolution: tables of plan PT built from that direction's twiddles
template <class A, class PT>
void fill_fc_side(FCSide<A>& S, const typename A::C* h_tw, const typename A::C* d_tw, const typename A::C* d_gtw)
{
    constexpr PlanDesc D = PT::D;
    KParams<A> P{};
    fill_g0tw<A, PT>(P, h_tw);
    for (int i = 0; i < kMaxG0Slots; ++i) S.g0tw[i] = P.g0tw[i];
    for (int i = 0; i < kMaxCtw; ++i) S.ctw[i] = P.ctw[i];
    S.tw = d_tw;
    S.gtw = d_gtw;
    const int N = D.N;
    typename A::C z{};
    S.pc.epi3 = A::load((N % 3 == 0) ? h_tw[N / 3] : z);
    S.pc.ya = A::load((N % 5 == 0) ? h_tw[N / 5] : z);
    S.pc.yb = A::load((N % 5 == 0) ? h_tw[2 * (N / 5)] : z);
}

// How many leading rows of a call may go through the fused kernel of plan PT (the rest, if any, must take the
// run-time kernel).  The bulk-async input ring needs contiguous rows, a 16-byte aligned base and tiles whose byte
// size is a multiple of 16 (cp.async.bulk rules); a ragged last tile that breaks the size rule is peeled off.
// Plans whose tiles are not a multiple of 16 bytes (FusedLayout::kSlackRing) copy from the aligned address below.
template <class A, class PT, int MODE>
long long fused_rows(const KParams<A>& P)
{
    typedef FusedLayout<A, PT, MODE> LY;
    constexpr PlanDesc D = PT::D;
    if (MODE == kC2C && P.in_stride != 1) return 0;
    if (!LY::kRing) return P.howmany;
    const size_t row_bytes = (size_t)LY::kRowIn * sizeof(typename A::C);
    if (P.in_dist != LY::kRowIn || ((size_t)P.in % 16) != 0) return 0;
    const long long rem = P.howmany % D.tpc;
    if (LY::kSlackRing) {
        // tiles start off the 16-byte grid and their copies are rounded up: the LAST tile may read up to 15 bytes past
        // the end of the batch, so it is kept only when the batch ends on the grid
        if (((size_t)P.howmany * row_bytes) % 16 == 0) return P.howmany;
        return P.howmany - (rem ? rem : D.tpc);
    }
    return ((size_t)rem * row_bytes) % 16 == 0 ? P.howmany : P.howmany - rem;
}

}   // namespace kf
