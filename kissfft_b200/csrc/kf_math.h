// kf_math.h -- per-datatype arithmetic and the radix-2/3/4/5/generic butterflies, written for registers.
//
// What this replaces in the reference (all citations relative to the kissfft source tree):
//   * the C_MUL / C_FIXDIV / S_MUL / HALF_OF / sround macro family   (_kiss_fft_guts.h:64-139)
//   * kf_bfly2 / kf_bfly3 / kf_bfly4 / kf_bfly5 / kf_bfly_generic      (kiss_fft.c:15-233)
//
// The reference butterflies walk an m-long loop over memory; here a butterfly is a pure function of p
// register-resident operands plus their twiddles, so that several consecutive radix stages can be chained
// in registers by one thread.  The fixed-point variants keep every rounding of the reference (one sround per
// S_MUL, one per C_MUL component after the exact sum of two products, the integer SAMP_MAX/p pre-scale, the
// arithmetic-shift HALF_OF) and therefore produce bit-identical Q15/Q31 results.
//
// Everything is usable from host code as well (KF_HD) so that the index math of the kernels can be
// exercised without a GPU by tests/emul (test-only; the product never runs these on the CPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KF_HD __host__ __device__ __forceinline__
#else
#define KF_HD inline __attribute__((always_inline))
#endif

namespace kf {

template <typename R>
struct cx {
    R r, i;
};

// ---------------------------------------------------------------------------------------------------------
// Arith<S>: S is the storage scalar (kiss_fft_scalar of the reference build being replaced).
//   R    register scalar (fixed point is widened to 32-bit registers)
//   C    storage complex as it lies in HBM / shared memory (same layout as kiss_fft_cpx, kiss_fft.h:87-90)
// ---------------------------------------------------------------------------------------------------------
template <typename S>
struct Arith;

template <typename F>
struct ArithFloat {
    typedef F S;
    typedef F R;
    struct alignas(2 * sizeof(F)) C {
        F r, i;
    };
    static constexpr bool kFixed = false;
    static KF_HD cx<R> load(const C& c) { return cx<R>{c.r, c.i}; }
    static KF_HD C store(const cx<R>& v) { return C{v.r, v.i}; }
    static KF_HD R add(R a, R b) { return a + b; }
    static KF_HD R sub(R a, R b) { return a - b; }
    static KF_HD R neg(R a) { return -a; }
    static KF_HD R mad_sign(R sg, R a, R b) { return sg * a + b; }   // sg == +-1: exact
    static KF_HD R sign_of(int inverse) { return inverse ? (F)-1 : (F)1; }
    static KF_HD R wrap(R a) { return a; }
    static KF_HD R smul(R a, R b) { return a * b; }                    // S_MUL, _kiss_fft_guts.h:86
    static KF_HD R half(R a) { return a * (F)0.5; }                     // HALF_OF, :138
    template <int K>
    static KF_HD R divk(R a) { return a; }                             // C_FIXDIV is a no-op, :90
    static KF_HD R divk_rt(R a, int) { return a; }
    static KF_HD cx<R> cmul(const cx<R>& a, const cx<R>& b)             // C_MUL, :87-89
    {
        return cx<R>{a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r};
    }
    static KF_HD cx<R> cmul_bf(const cx<R>& a, const cx<R>& b) { return cmul(a, b); }
    static KF_HD R smul_bf(R a, R b) { return a * b; }
};
template <>
struct Arith<float> : ArithFloat<float> {};
template <>
struct Arith<double> : ArithFloat<double> {};

// Q15: int16_t storage, int32 products (FRACBITS 15, SAMP_MAX 32767; _kiss_fft_guts.h:50-55)
template <>
struct Arith<int16_t> {
    typedef int16_t S;
    typedef int32_t R;
    struct alignas(4) C {
        int16_t r, i;
    };
    static constexpr bool kFixed = true;
    static constexpr int kFrac = 15;
    static constexpr int32_t kSampMax = 32767;
    static KF_HD cx<R> load(const C& c) { return cx<R>{(R)c.r, (R)c.i}; }
    static KF_HD C store(const cx<R>& v) { return C{(int16_t)v.r, (int16_t)v.i}; }
    // C_ADD & friends store back into int16_t (:100-124): wrap-around on store.  Adds are done unsigned so the
    // compiler may not assume "no overflow"; wrap() re-creates the int16 truncation where a value is consumed
    // by a multiply or a shift.
    static KF_HD R add(R a, R b) { return (R)((uint32_t)a + (uint32_t)b); }
    static KF_HD R sub(R a, R b) { return (R)((uint32_t)a - (uint32_t)b); }
    static KF_HD R neg(R a) { return (R)(0u - (uint32_t)a); }
    static KF_HD R mad_sign(R sg, R a, R b) { return (R)((uint32_t)sg * (uint32_t)a + (uint32_t)b); }
    static KF_HD R sign_of(int inverse) { return inverse ? -1 : 1; }
    static KF_HD R wrap(R a) { return (R)(int16_t)a; }
    static KF_HD R sround(int32_t x) { return (R)(int16_t)((x + (1 << (kFrac - 1))) >> kFrac); }   // :65
    // same value without re-creating the int16 truncation: only where the result provably fits in int16 --
    //   * x * (SAMP_MAX/k), k >= 2, |x| <= 32768                       -> |result| <= 16384
    //   * a * twiddle inside a butterfly, a pre-scaled by 1/p (p >= 2)  -> |result| <= 23171 (|a| <= 16384*sqrt2, |tw| ~ 32767)
    //   * S_MUL of a butterfly-internal sum with a twiddle component    -> |result| <= 30894
    static KF_HD R sround_fit(int32_t x) { return (R)((x + (1 << (kFrac - 1))) >> kFrac); }
    static KF_HD R smul(R a, R b) { return sround(a * b); }                                        // :67
    static KF_HD R smul_bf(R a, R b) { return sround_fit(a * b); }
    static KF_HD R half(R a) { return a >> 1; }                                                    // :130
    template <int K>
    static KF_HD R divk(R a) { return sround_fit(a * (kSampMax / K)); }                            // :73-78
    static KF_HD R divk_rt(R a, int k) { return sround_fit(a * (kSampMax / k)); }
    static KF_HD cx<R> cmul(const cx<R>& a, const cx<R>& b)                                        // :69-71
    {
        return cx<R>{sround(a.r * b.r - a.i * b.i), sround(a.r * b.i + a.i * b.r)};
    }
    static KF_HD cx<R> cmul_bf(const cx<R>& a, const cx<R>& b)
    {
        return cx<R>{sround_fit(a.r * b.r - a.i * b.i), sround_fit(a.r * b.i + a.i * b.r)};
    }
};

// Q31: int32_t storage, int64 products (FRACBITS 31, SAMP_MAX 2^31-1; _kiss_fft_guts.h:45-49)
//
// On the device every product-sum is a chain of mad.wide.s32 (IMAD.WIDE: 32 x 32 + 64 -> 64 in one instruction) seeded with
// the rounding constant 2^30, and "(x >> 31) truncated to 32 bits" is one funnel shift of the accumulator's two halves.
// Written as plain int64 C the compiler narrows the expression to the bits it needs and rebuilds it from 32-bit
// mul.lo / mul.hi pieces with carry chains -- 2.9 k instructions per 16-point work item instead of ~1.5 k (round 2,
// profiles/r02/ncu_all_kernels.txt: 622 IMAD + 752 IADD3/.X + 420 SHF around 466 wide multiplies).  Same integers either way.
template <>
struct Arith<int32_t> {
    typedef int32_t S;
    typedef int32_t R;
    struct alignas(8) C {
        int32_t r, i;
    };
    static constexpr bool kFixed = true;
    static constexpr int kFrac = 31;
    static constexpr int32_t kSampMax = 2147483647;
    static constexpr int64_t kRound = (int64_t)1 << (kFrac - 1);
    static KF_HD cx<R> load(const C& c) { return cx<R>{c.r, c.i}; }
    static KF_HD C store(const cx<R>& v) { return C{v.r, v.i}; }
    static KF_HD R add(R a, R b) { return (R)((uint32_t)a + (uint32_t)b); }
    static KF_HD R sub(R a, R b) { return (R)((uint32_t)a - (uint32_t)b); }
    static KF_HD R neg(R a) { return (R)(0u - (uint32_t)a); }
    static KF_HD R mad_sign(R sg, R a, R b) { return (R)((uint32_t)sg * (uint32_t)a + (uint32_t)b); }
    static KF_HD R sign_of(int inverse) { return inverse ? -1 : 1; }
    static KF_HD R wrap(R a) { return a; }
    // a*b + c in 64 bits
    static KF_HD int64_t madw(int32_t a, int32_t b, int64_t c)
    {
#if defined(__CUDA_ARCH__)
        int64_t d;
        asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
        return d;
#else
        return (int64_t)a * b + c;
#endif
    }
    // low 32 bits of x >> 31 (the rounding constant is already in x): sround, _kiss_fft_guts.h:65
    static KF_HD R shr31(int64_t x)
    {
#if defined(__CUDA_ARCH__)
        unsigned lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
        return (R)__funnelshift_l(lo, hi, 1);
#else
        return (R)(x >> kFrac);
#endif
    }
    static KF_HD R sround(int64_t x) { return shr31(x + kRound); }
    static KF_HD R smul(R a, R b) { return shr31(madw(a, b, kRound)); }                           // S_MUL, :67
    static KF_HD R half(R a) { return a >> 1; }
    // C_FIXDIV (:73-78): (a*c + 2^30) >> 31 with c = SAMP_MAX/K, < 2^30 for every K >= 2.  Doubling the constant moves the wanted bits into
    // the upper word: (a*2c + 2^31) >> 32 = hi + (lo >> 31) -- one wide multiply and one shift-add, no 64-bit addition
    static KF_HD R divc(R a, int32_t c)
    {
#if defined(__CUDA_ARCH__)
        if (c >= (1 << 30)) return shr31(madw(a, c, kRound));      // K == 1 (nfft == 1, kf_bfly_generic with p == 1): 2c does not fit
        unsigned lo, hi;
        asm("{\n\t.reg .b64 t;\n\tmul.wide.s32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(2 * c));
        return (R)(hi + (lo >> 31));
#else
        return shr31(madw(a, c, kRound));
#endif
    }
    template <int K>
    static KF_HD R divk(R a) { return divc(a, kSampMax / K); }
    static KF_HD R divk_rt(R a, int k) { return divc(a, kSampMax / k); }
    // C_MUL (:69-71); b is always the table operand (a twiddle, |b| <= SAMP_MAX), so negating b.i cannot overflow and
    // a.r*b.r + a.i*(-b.i) is the reference's a.r*b.r - a.i*b.i in exact 64-bit arithmetic
    static KF_HD cx<R> cmul(const cx<R>& a, const cx<R>& b)
    {
        return cx<R>{shr31(madw(a.i, (R)(0 - b.i), madw(a.r, b.r, kRound))), shr31(madw(a.i, b.r, madw(a.r, b.i, kRound)))};
    }
    static KF_HD cx<R> cmul_bf(const cx<R>& a, const cx<R>& b) { return cmul(a, b); }
    static KF_HD R smul_bf(R a, R b) { return smul(a, b); }
};

template <class A>
KF_HD cx<typename A::R> cadd(const cx<typename A::R>& a, const cx<typename A::R>& b)
{
    return cx<typename A::R>{A::add(a.r, b.r), A::add(a.i, b.i)};
}
template <class A>
KF_HD cx<typename A::R> csub(const cx<typename A::R>& a, const cx<typename A::R>& b)
{
    return cx<typename A::R>{A::sub(a.r, b.r), A::sub(a.i, b.i)};
}
template <class A, int K>
KF_HD cx<typename A::R> cfixdiv(const cx<typename A::R>& a)
{
    return cx<typename A::R>{A::template divk<K>(a.r), A::template divk<K>(a.i)};
}
template <class A>
KF_HD cx<typename A::R> cwrap(const cx<typename A::R>& a)
{
    return cx<typename A::R>{A::wrap(a.r), A::wrap(a.i)};
}

// ---------------------------------------------------------------------------------------------------------
// Butterflies.  Operands v[q] are the p inputs Y_{s+1}[off + q*F_s][k]; on return v[r] = Y_s[off][k + r*m_s].
// Inputs must be "wrapped" (valid scalars); outputs are wrapped again so stages can be chained in registers.
// TW1: when true the twiddles are known to be exactly (1,0) and the products are skipped -- only legal for
// float/double (there tw[0] == (1,0) exactly; in fixed point tw[0] == (SAMP_MAX,0) changes values and the
// reference never skips it, kiss_fft.c:57-59).
// ---------------------------------------------------------------------------------------------------------

// kf_bfly2, kiss_fft.c:15-36
template <class A, bool TW1>
KF_HD void bfly2(cx<typename A::R>* v, const cx<typename A::R>& t1)
{
    typedef cx<typename A::R> X;
    X a = cfixdiv<A, 2>(v[0]), b = cfixdiv<A, 2>(v[1]);
    X t = TW1 ? b : A::cmul_bf(b, t1);
    v[1] = cwrap<A>(csub<A>(a, t));
    v[0] = cwrap<A>(cadd<A>(a, t));
}

// kf_bfly4, kiss_fft.c:38-84
template <class A, bool TW1>
KF_HD void bfly4(cx<typename A::R>* v, const cx<typename A::R>& t1, const cx<typename A::R>& t2,
                 const cx<typename A::R>& t3, typename A::R sg)
{
    typedef cx<typename A::R> X;
    X f0 = cfixdiv<A, 4>(v[0]), f1 = cfixdiv<A, 4>(v[1]), f2 = cfixdiv<A, 4>(v[2]), f3 = cfixdiv<A, 4>(v[3]);
    X s0 = TW1 ? f1 : A::cmul_bf(f1, t1);
    X s1 = TW1 ? f2 : A::cmul_bf(f2, t2);
    X s2 = TW1 ? f3 : A::cmul_bf(f3, t3);
    X s5 = csub<A>(f0, s1);
    f0 = cadd<A>(f0, s1);
    X s3 = cadd<A>(s0, s2);
    X s4 = csub<A>(s0, s2);
    v[2] = cwrap<A>(csub<A>(f0, s3));
    v[0] = cwrap<A>(cadd<A>(f0, s3));
    // forward: F[k+m] = s5 - j*s4, F[k+3m] = s5 + j*s4; inverse swaps them (kiss_fft.c:71-81).  The direction
    // enters as sg = +1 (forward) / -1 (inverse) through a multiply-add, which is exact (|sg| == 1) and costs
    // the same instruction as the add it replaces -- no select, no second kernel instantiation.
    v[1] = cwrap<A>(X{A::mad_sign(sg, s4.i, s5.r), A::mad_sign(A::neg(sg), s4.r, s5.i)});
    v[3] = cwrap<A>(X{A::mad_sign(A::neg(sg), s4.i, s5.r), A::mad_sign(sg, s4.r, s5.i)});
}

// kf_bfly3, kiss_fft.c:86-128.  epi3i = twiddles[N/3].i (sign carries the direction)
template <class A, bool TW1>
KF_HD void bfly3(cx<typename A::R>* v, const cx<typename A::R>& t1, const cx<typename A::R>& t2, typename A::R epi3i)
{
    typedef cx<typename A::R> X;
    X f0 = cfixdiv<A, 3>(v[0]), f1 = cfixdiv<A, 3>(v[1]), f2 = cfixdiv<A, 3>(v[2]);
    X s1 = TW1 ? f1 : A::cmul_bf(f1, t1);
    X s2 = TW1 ? f2 : A::cmul_bf(f2, t2);
    X s3 = cwrap<A>(cadd<A>(s1, s2));
    X s0 = cwrap<A>(csub<A>(s1, s2));
    X fm{A::sub(f0.r, A::half(s3.r)), A::sub(f0.i, A::half(s3.i))};
    s0.r = A::smul_bf(s0.r, epi3i);   // C_MULBYSCALAR
    s0.i = A::smul_bf(s0.i, epi3i);
    v[0] = cwrap<A>(cadd<A>(f0, s3));
    v[2] = cwrap<A>(X{A::add(fm.r, s0.i), A::sub(fm.i, s0.r)});
    v[1] = cwrap<A>(X{A::sub(fm.r, s0.i), A::add(fm.i, s0.r)});
}

// kf_bfly5, kiss_fft.c:130-189.  ya = twiddles[N/5], yb = twiddles[2N/5]
template <class A, bool TW1>
KF_HD void bfly5(cx<typename A::R>* v, const cx<typename A::R>& t1, const cx<typename A::R>& t2,
                 const cx<typename A::R>& t3, const cx<typename A::R>& t4, const cx<typename A::R>& ya,
                 const cx<typename A::R>& yb)
{
    typedef typename A::R R;
    typedef cx<R> X;
    X f0 = cfixdiv<A, 5>(v[0]), f1 = cfixdiv<A, 5>(v[1]), f2 = cfixdiv<A, 5>(v[2]), f3 = cfixdiv<A, 5>(v[3]),
      f4 = cfixdiv<A, 5>(v[4]);
    X s1 = TW1 ? f1 : A::cmul_bf(f1, t1);
    X s2 = TW1 ? f2 : A::cmul_bf(f2, t2);
    X s3 = TW1 ? f3 : A::cmul_bf(f3, t3);
    X s4 = TW1 ? f4 : A::cmul_bf(f4, t4);
    X s7 = cwrap<A>(cadd<A>(s1, s4)), s10 = cwrap<A>(csub<A>(s1, s4));
    X s8 = cwrap<A>(cadd<A>(s2, s3)), s9 = cwrap<A>(csub<A>(s2, s3));

    v[0] = cwrap<A>(X{A::add(f0.r, A::add(s7.r, s8.r)), A::add(f0.i, A::add(s7.i, s8.i))});

    X s5{A::add(A::add(f0.r, A::smul_bf(s7.r, ya.r)), A::smul_bf(s8.r, yb.r)),
         A::add(A::add(f0.i, A::smul_bf(s7.i, ya.r)), A::smul_bf(s8.i, yb.r))};
    X s6{A::add(A::smul_bf(s10.i, ya.i), A::smul_bf(s9.i, yb.i)),
         A::sub(A::neg(A::smul_bf(s10.r, ya.i)), A::smul_bf(s9.r, yb.i))};
    v[1] = cwrap<A>(csub<A>(s5, s6));
    v[4] = cwrap<A>(cadd<A>(s5, s6));

    X s11{A::add(A::add(f0.r, A::smul_bf(s7.r, yb.r)), A::smul_bf(s8.r, ya.r)),
          A::add(A::add(f0.i, A::smul_bf(s7.i, yb.r)), A::smul_bf(s8.i, ya.r))};
    X s12{A::add(A::neg(A::smul_bf(s10.i, yb.i)), A::smul_bf(s9.i, ya.i)),
          A::sub(A::smul_bf(s10.r, yb.i), A::smul_bf(s9.r, ya.i))};
    v[2] = cwrap<A>(cadd<A>(s11, s12));
    v[3] = cwrap<A>(csub<A>(s11, s12));
}

}   // namespace kf
