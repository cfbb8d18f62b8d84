// kf_engine.h -- register-group executor of a compile-time Stockham plan (see kf_plan.h for the math).
//
// Replaces kf_work's depth-first recursion (kiss_fft.c:235-300): one "work item" of group g is the set of
// R(g) elements that the fused stages s_hi(g)..s_lo(g) combine; a thread loads them (from HBM for the first
// group, from the shared-memory exchange buffer otherwise), runs the fused radix stages in registers with the
// butterflies of kf_math.h, and stores them (to HBM for the last group).  All index arithmetic that depends
// only on the plan is resolved at compile time (static_for + constexpr PlanDesc members).
//
// Host-compilable (KF_HD) so tests/emul can execute the same index math thread-by-thread on the CPU.
#pragma once
#include <type_traits>
#include <utility>

#include "kf_plan.h"

namespace kf {

template <class F, int... I>
KF_HD void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
KF_HD void static_for(F&& f)
{
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// read-only table fetch (twiddles live in HBM but are L1/L2 resident: <= 32 KiB per plan)
template <class T>
KF_HD T ro_load(const T* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <class A>
struct TwTab {
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    const C* tw;    // N entries, tw[i] = exp(-+2 pi j i/N) generated on the host exactly like kiss_fft.c:361-367
    const C* gtw;   // per-group stage-twiddle tables [g][slot][work item] (PlanDesc::slot), exact copies of tw[] entries
    const X* g0;    // group 0's slots (kernel parameter space)
    const X* ctw;   // split-twiddle constants (kernel parameter space), PlanDesc::cslot
    const X* hreg;  // this thread's loop-invariant stage twiddles (registers), PlanDesc::hoist_base; null when not hoisted
    KF_HD X get(int idx) const
    {
        // all storage complexes are plain {S r, i}; load as one vector
        C c = ro_load_c(tw + idx);
        return A::load(c);
    }
    static KF_HD C ro_load_c(const C* p)
    {
#if defined(__CUDA_ARCH__)
        if constexpr (sizeof(C) == 4) {
            unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
            return *reinterpret_cast<C*>(&u);
        } else if constexpr (sizeof(C) == 8) {
            float2 u = __ldg(reinterpret_cast<const float2*>(p));
            return *reinterpret_cast<C*>(&u);
        } else {
            double2 u = __ldg(reinterpret_cast<const double2*>(p));
            return *reinterpret_cast<C*>(&u);
        }
#else
        return *p;
#endif
    }
};

// per-plan constants used by the radix-3 and radix-5 butterflies (kiss_fft.c:99, 143-144)
template <class A>
struct PlanConsts {
    cx<typename A::R> epi3;   // tw[N/3]
    cx<typename A::R> ya;     // tw[N/5]
    cx<typename A::R> yb;     // tw[2N/5]
};

// stage twiddle of slot `slot` (compile-time) for work item w of group g.  HB >= 0: the thread fetched this item's
// twiddles into registers before its tile loop (fused_body, PlanDesc::hoist) and HB is their base index in tw.hreg.
template <class A, PlanDesc D, int g, int HB = -1>
KF_HD cx<typename A::R> stage_tw(const TwTab<A>& tw, int slot, int w)
{
    if constexpr (g == 0) return tw.g0[slot];
    else if constexpr (HB >= 0) return tw.hreg[HB + slot];
    else return A::load(TwTab<A>::ro_load_c(tw.gtw + (D.gtw_offset(g) + slot * D.items(g)) + w));
}

// fetches the stage twiddles of work item w of group g (the slots run_stage reads from the table) into hreg[HB ...]
template <class A, PlanDesc D, int g, int HB>
KF_HD void hoist_item(const TwTab<A>& tw, int w, cx<typename A::R>* hreg)
{
    static_for<D.glen[g]>([&](auto SS) {
        constexpr int s = D.s_hi(g) - decltype(SS)::value;
        static_for<D.R(g)>([&](auto E) {
            constexpr int e = decltype(E)::value;
            if constexpr (D.slot_fetched(g, s, e, A::kFixed)) {
                static_for<D.p[s] - 1>([&](auto Q) {
                    constexpr int slot = D.slot(g, s, e) + decltype(Q)::value;
                    hreg[HB + slot] = stage_tw<A, D, g>(tw, slot, w);
                });
            }
        });
    });
}

// ---------------------------------------------------------------------------------------------------------
// generic radix p (kf_bfly_generic, kiss_fft.c:192-233)
// ---------------------------------------------------------------------------------------------------------
// Fixed point: literal restatement -- scratch pre-scaled by 1/p, twiddle index walked modulo N, products
// rounded one by one and accumulated in the order q = 1..p-1, so the result is bit-identical.
template <class A, int N, int p, int Fs, int ms>
KF_HD void bfly_generic_fixed(cx<typename A::R>* v, int ks, const TwTab<A>& tw)
{
    typedef cx<typename A::R> X;
    X sc[p];
    static_for<p>([&](auto Q) { constexpr int q = decltype(Q)::value; sc[q] = cfixdiv<A, p>(v[q]); });
    static_for<p>([&](auto Q1) {
        constexpr int q1 = decltype(Q1)::value;
        const int k = ks + q1 * ms;
        int twidx = 0;
        X acc = sc[0];
        static_for<p - 1>([&](auto Qm1) {
            twidx += Fs * k;
            if (twidx >= N) twidx -= N;
            acc = cadd<A>(acc, A::cmul_bf(sc[decltype(Qm1)::value + 1], tw.get(twidx)));
        });
        v[q1] = cwrap<A>(acc);
    });
}

// Floating point: tw[(q*(ks+q1*ms)*Fs) mod N] = tw[q*ks*Fs] * W_p^(q*q1) with W_p^j = tw[j*N/p], so the stage
// twiddles are applied first (p-1 products, as the radix-2..5 butterflies do) and the remaining constant
// p-point DFT uses the conjugate symmetry W_p^(p-j) = conj(W_p^j): outputs q1 and p-q1 share their sums.
// Same operands, different association => equal up to rounding (parity is relative-RMS for float/double).
template <class A, PlanDesc D, int g, int p, bool TW1, int HB = -1>
KF_HD void bfly_generic_float(cx<typename A::R>* v, int slot0, int w, const TwTab<A>& tw)
{
    constexpr int N = D.N;
    typedef typename A::R R;
    typedef cx<R> X;
    static_assert(p % 2 == 1, "generic radices are odd (kf_factor emits 4, 2, then odd numbers)");
    constexpr int h = (p - 1) / 2;
    X y[p];
    y[0] = v[0];
    static_for<p - 1>([&](auto Qm1) {
        constexpr int q = decltype(Qm1)::value + 1;
        y[q] = TW1 ? v[q] : A::cmul_bf(v[q], stage_tw<A, D, g, HB>(tw, slot0 + q - 1, w));
    });
    X a[h + 1], b[h + 1];
    X sum = y[0];
    static_for<h>([&](auto Qm1) {
        constexpr int q = decltype(Qm1)::value + 1;
        a[q] = X{y[q].r + y[p - q].r, y[q].i + y[p - q].i};
        b[q] = X{y[q].r - y[p - q].r, y[q].i - y[p - q].i};
        sum.r += a[q].r;
        sum.i += a[q].i;
    });
    v[0] = sum;
    static_for<h>([&](auto Q1m1) {
        constexpr int q1 = decltype(Q1m1)::value + 1;
        X P = y[0], Qv{(R)0, (R)0};
        static_for<h>([&](auto Qm1) {
            constexpr int q = decltype(Qm1)::value + 1;
            constexpr int j = (q * q1) % p;
            const X w = tw.get(j * (N / p));   // compile-time index: uniform, cache resident
            P.r += a[q].r * w.r;
            P.i += a[q].i * w.r;
            Qv.r -= b[q].i * w.i;
            Qv.i += b[q].r * w.i;
        });
        v[q1] = X{P.r + Qv.r, P.i + Qv.i};
        v[p - q1] = X{P.r - Qv.r, P.i - Qv.i};
    });
}

// ---------------------------------------------------------------------------------------------------------
// one radix stage s of group g applied to the R(g) registers of a work item whose k' is `kp`
// ---------------------------------------------------------------------------------------------------------
template <class A, PlanDesc D, int g, int s, int HB = -1>
KF_HD void run_stage(cx<typename A::R>* v, int kp, int w, const TwTab<A>& tw, const PlanConsts<A>& pc, typename A::R sg)
{
    typedef cx<typename A::R> X;
    constexpr int p = D.p[s], Ws = D.W(g, s), Fs = D.F(s), ms = D.m(s), R = D.R(g);
    static_for<R>([&](auto E) {
        constexpr int e = decltype(E)::value;
        if constexpr (D.digit(g, s, e) == 0) {
            constexpr int kab = D.kabove(g, s, e);
            // group 0 has k' == 0 (m_{L-1} == 1), so its twiddle indices are compile-time constants and the
            // all-ones case can be dropped for float/double
            constexpr bool kTw1 = !A::kFixed && g == 0 && kab == 0;
            const int ks = (g == 0 ? 0 : kp) + kab;
            X x[p];
            static_for<p>([&](auto Q) { constexpr int q = decltype(Q)::value; x[q] = v[e + q * Ws]; });
            constexpr int sl = D.slot(g, s, e);
            constexpr int up = e / (Ws * p);
            // split mode: twiddle(q) = [table entry of the upper == 0 butterfly of this stage] * [plan constant]
            constexpr bool kSplit = !A::kFixed && D.twmode == 1 && g > 0 && up > 0;
            auto T = [&](int q) {
                if constexpr (kSplit) {
                    constexpr int sl0 = D.slot(g, s, e - D.upper_base(g, s, up));
                    return A::cmul(stage_tw<A, D, g, HB>(tw, sl0 + q - 1, w), tw.ctw[D.cslot(g, s, up, 1) + q - 1]);
                } else {
                    return stage_tw<A, D, g, HB>(tw, sl + q - 1, w);
                }
            };
            if constexpr (p == 2) {
                bfly2<A, kTw1>(x, kTw1 ? X{} : T(1));
            } else if constexpr (p == 4) {
                bfly4<A, kTw1>(x, kTw1 ? X{} : T(1), kTw1 ? X{} : T(2), kTw1 ? X{} : T(3), sg);
            } else if constexpr (p == 3) {
                bfly3<A, kTw1>(x, kTw1 ? X{} : T(1), kTw1 ? X{} : T(2), pc.epi3.i);
            } else if constexpr (p == 5) {
                bfly5<A, kTw1>(x, kTw1 ? X{} : T(1), kTw1 ? X{} : T(2), kTw1 ? X{} : T(3), kTw1 ? X{} : T(4), pc.ya, pc.yb);
            } else if constexpr (A::kFixed) {
                bfly_generic_fixed<A, D.N, p, Fs, ms>(x, ks, tw);
            } else {
                bfly_generic_float<A, D, g, p, kTw1, HB>(x, sl, w, tw);
            }
            static_for<p>([&](auto Q) { constexpr int q = decltype(Q)::value; v[e + q * Ws] = x[q]; });
        }
    });
}

template <class A, PlanDesc D, int g, int s, int HB = -1>
KF_HD void run_stages_from(cx<typename A::R>* v, int kp, int w, const TwTab<A>& tw, const PlanConsts<A>& pc, typename A::R sg)
{
    run_stage<A, D, g, s, HB>(v, kp, w, tw, pc, sg);
    if constexpr (s > D.s_lo(g)) run_stages_from<A, D, g, s - 1, HB>(v, kp, w, tw, pc, sg);
}

KF_HD int phys_rt(int a, int logpad) { return logpad >= 31 ? a : a + (a >> logpad); }

// ---------------------------------------------------------------------------------------------------------
// one group for one thread.  t = thread index inside the team, `active` = this team has a transform.
//   Src::load(i)      element i of the level-L (natural order) input of this transform   [group 0 only]
//   Dst::store(k, v)  element k of the natural-order output                               [last group only]
//   rd / wr           this transform's exchange buffers (already offset by team * pitch)
// ---------------------------------------------------------------------------------------------------------
// the R(g) registers of work item w of group g >= 1, read from the exchange buffer the previous group wrote
template <class A, PlanDesc D, int g>
KF_HD void item_load(int w, const typename A::C* rd, cx<typename A::R>* v)
{
    constexpr int R = D.R(g), Flo = D.Flo(g);
    // When the skew phys(a) = a + (a >> logpad) cannot carry between the thread-dependent base and the
    // register-dependent offset (true for all power-of-two plans, see PlanDesc::lin_rd/lin_wr) the offset part is
    // a compile-time constant and folds into the LDS/STS immediate.
    constexpr bool kLinRd = D.lin_rd(g);
    const int off = w % Flo, kp = w / Flo;
    const int rbase = phys_rt(kp * (Flo * R) + off, D.logpad);
    static_for<R>([&](auto E) {
        constexpr int e = decltype(E)::value;
        if constexpr (kLinRd) v[e] = A::load(rd[rbase + D.phys(e * Flo)]);
        else v[e] = A::load(rd[phys_rt(kp * (Flo * R) + off + e * Flo, D.logpad)]);
    });
}

// the radix stages of group g on the registers of work item w
template <class A, PlanDesc D, int g, int HB = -1>
KF_HD void item_stages(int w, cx<typename A::R>* v, const TwTab<A>& tw, const PlanConsts<A>& pc, typename A::R sg)
{
    run_stages_from<A, D, g, D.s_hi(g), HB>(v, w / D.Flo(g), w, tw, pc, sg);
}

// the outputs of work item w of group g < G-1, written to the exchange buffer the next group reads
template <class A, PlanDesc D, int g>
KF_HD void item_store(int w, typename A::C* wr, const cx<typename A::R>* v)
{
    constexpr int R = D.R(g), Flo = D.Flo(g);
    constexpr bool kLinWr = D.lin_wr(g);
    const int off = w % Flo, kp = w / Flo;
    const int wbase = phys_rt(kp * Flo + off, D.logpad);
    static_for<R>([&](auto E) {
        constexpr int e = decltype(E)::value;
        if constexpr (kLinWr) wr[wbase + D.phys(D.kout(g, e) * Flo)] = A::store(v[e]);
        else wr[phys_rt((kp + D.kout(g, e)) * Flo + off, D.logpad)] = A::store(v[e]);
    });
}

template <class A, PlanDesc D, int g, class Src, class Dst>
KF_HD void run_group(int t, bool active, const Src& src, const Dst& dst, const typename A::C* rd, typename A::C* wr,
                     const TwTab<A>& tw, const PlanConsts<A>& pc, int inverse)
{
    typedef cx<typename A::R> X;
    constexpr int R = D.R(g), WI = D.items(g), IT = D.iters(g), Flo = D.Flo(g);
    constexpr bool kFirst = (g == 0), kLast = (g == D.G - 1);
    const typename A::R sg = A::sign_of(inverse);
    static_for<IT>([&](auto ITER) {
        constexpr int it = decltype(ITER)::value;
        const int w = t + it * D.team;
        const bool on = active && ((it + 1) * D.team <= WI || w < WI);
        if (on) {
            X v[R];
            if constexpr (kFirst) {
                static_for<R>([&](auto E) {
                    constexpr int e = decltype(E)::value;
                    v[e] = src.template get<it, e>(w % Flo + e * Flo);
                });
            } else {
                item_load<A, D, g>(w, rd, v);
            }
            // hoisted stage twiddles (ordinary groups >= 1 only; the item a thread runs in iteration `it` never changes)
            constexpr int kHB = ((D.hoist & 2) && g >= 1) ? D.hoist_base(g, it) : -1;
            item_stages<A, D, g, kHB>(w, v, tw, pc, sg);
            if constexpr (kLast) {
                static_for<R>([&](auto E) {
                    constexpr int e = decltype(E)::value;
                    dst.template put<it, e>(w / Flo + D.kout(g, e), v[e]);
                });
            } else {
                item_store<A, D, g>(w, wr, v);
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------------------
// real-transform split passes (kiss_fftr.c:88-116 and :131-153) on one bin pair k / ncfft-k
// ---------------------------------------------------------------------------------------------------------
// forward: T = FFT_ncfft(packed input); produces out[k] and out[nc-k] (k >= 1), or out[0] / out[nc] (k == 0)
template <class A>
KF_HD void fftr_post_pair(int k, int nc, const cx<typename A::R>& Tk, const cx<typename A::R>& Tnk,
                          const cx<typename A::R>& st_km1, cx<typename A::R>& outk, cx<typename A::R>& outnk)
{
    typedef typename A::R R;
    typedef cx<R> X;
    if (k == 0) {
        X tdc = cfixdiv<A, 2>(Tk);
        outk = X{A::wrap(A::add(tdc.r, tdc.i)), (R)0};
        outnk = X{A::wrap(A::sub(tdc.r, tdc.i)), (R)0};
        return;
    }
    X fpk = cfixdiv<A, 2>(Tk);
    X fpnk = cfixdiv<A, 2>(X{Tnk.r, A::wrap(A::neg(Tnk.i))});
    X f1k = cwrap<A>(cadd<A>(fpk, fpnk));
    X f2k = cwrap<A>(csub<A>(fpk, fpnk));
    X tw = A::cmul(f2k, st_km1);
    // HALF_OF(sum): the sum is formed in int before the shift (kiss_fftr.c:112-115)
    outk = X{A::half(A::add(f1k.r, tw.r)), A::half(A::add(f1k.i, tw.i))};
    outnk = X{A::half(A::sub(f1k.r, tw.r)), A::half(A::sub(tw.i, f1k.i))};
    (void)nc;
}

// inverse: consumes F[k], F[nc-k]; produces T[k], T[nc-k] (k >= 1) or T[0] (k == 0, Tnk_out unused)
template <class A>
KF_HD void fftri_pre_pair(int k, const cx<typename A::R>& Fk, const cx<typename A::R>& Fnk,
                          const cx<typename A::R>& st_km1, cx<typename A::R>& Tk, cx<typename A::R>& Tnk)
{
    typedef cx<typename A::R> X;
    if (k == 0) {
        // Fnk is F[nc] here
        X t{A::wrap(A::add(Fk.r, Fnk.r)), A::wrap(A::sub(Fk.r, Fnk.r))};
        Tk = cfixdiv<A, 2>(t);
        Tnk = Tk;
        return;
    }
    X fk = cfixdiv<A, 2>(Fk);
    X fnkc = cfixdiv<A, 2>(X{Fnk.r, A::wrap(A::neg(Fnk.i))});
    X fek = cwrap<A>(cadd<A>(fk, fnkc));
    X tmp = cwrap<A>(csub<A>(fk, fnkc));
    X fok = A::cmul(tmp, st_km1);
    Tk = cwrap<A>(cadd<A>(fek, fok));
    X d = csub<A>(fek, fok);
    Tnk = cwrap<A>(X{d.r, A::neg(d.i)});
}

}   // namespace kf
