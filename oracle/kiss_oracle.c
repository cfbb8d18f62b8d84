/*
 * kiss_oracle.c -- CPU restatement of KISS FFT's transform path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The product path (kissfft_b200/csrc) never
 * links, loads or calls this file.
 *
 * Parity: PINNED.  tests/test_oracle_pin.py compares every entry point below bit-for-bit
 * (Q15/Q31, and in practice float/double too) against the reference library compiled from
 * /root/reference by oracle/Makefile (oracle/_ref/libkissfft-<type>.so) and against the golden
 * vectors under tests/golden/ that were generated from that compiled reference.
 *
 * This is NOT a copy of the reference: the reference is a depth-first recursion that
 * digit-reverses on the way down and runs in-place butterflies on the way up
 * (kiss_fft.c:235-300).  This file restates the same arithmetic as an iterative, breadth-first
 * sequence of whole-array passes in the "layout-free" form
 *
 *     Y_s[off][k + r*m_s] = butterfly_{p_s}( Y_{s+1}[off + q*F_s][k], q = 0..p_s-1 ),
 *     F_s = prod_{j<s} p_j,   Y_L[off][0] = x[off],   X[k] = Y_0[0][k]
 *
 * which is exactly the formulation the CUDA kernels implement (autosort addressing
 * addr_s(off,k) = k*F_s + off), so the operands of every butterfly are the reference's operands
 * even though the memory order differs.
 *
 * One shared library per datatype, selected exactly like the reference (kiss_fft.h:73-85):
 *   -DFIXED_POINT=16 | -DFIXED_POINT=32 | -Dkiss_fft_scalar=double | (default float)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef FIXED_POINT
# if (FIXED_POINT == 32)
typedef int32_t osc;
typedef int64_t oprod;
#  define O_FRACBITS 31
#  define O_SAMP_MAX INT32_MAX
# else
typedef int16_t osc;
typedef int32_t oprod;
#  define O_FRACBITS 15
#  define O_SAMP_MAX INT16_MAX
# endif
#else
# ifndef kiss_fft_scalar
#  define kiss_fft_scalar float
# endif
typedef kiss_fft_scalar osc;
#endif

typedef struct { osc r, i; } ocpx;

#define O_MAXSTAGES 32 /* _kiss_fft_guts.h:21 MAXFACTORS */

/* ---- scalar primitives: _kiss_fft_guts.h:64-93, 127-139 ------------------------------------ */
#ifdef FIXED_POINT
/* sround: add half an LSB of the product, arithmetic shift right, truncate to the scalar type
 * (_kiss_fft_guts.h:65) */
static inline osc o_round(oprod x) { return (osc)((x + ((oprod)1 << (O_FRACBITS - 1))) >> O_FRACBITS); }
static inline osc o_smul(osc a, osc b) { return o_round((oprod)a * b); }            /* S_MUL, :67 */
static inline osc o_divk(osc x, int k) { return o_round((oprod)x * (O_SAMP_MAX / k)); } /* DIVSCALAR, :73-74 */
static inline osc o_half(osc x) { return (osc)(x >> 1); }                           /* HALF_OF, :130 */
static inline osc o_half_sum(osc a, osc b) { return (osc)(((int)a + (int)b) >> 1); } /* HALF_OF(a+b) as used in kiss_fftr.c:112-115 */
static inline osc o_half_dif(osc a, osc b) { return (osc)(((int)a - (int)b) >> 1); }
static inline ocpx o_cmul(ocpx a, ocpx b)                                           /* C_MUL, :69-71 */
{
    ocpx m;
    m.r = o_round((oprod)a.r * b.r - (oprod)a.i * b.i);
    m.i = o_round((oprod)a.r * b.i + (oprod)a.i * b.r);
    return m;
}
static inline ocpx o_fixdiv(ocpx c, int k) { c.r = o_divk(c.r, k); c.i = o_divk(c.i, k); return c; } /* C_FIXDIV, :76-78 */
static inline osc o_cos(double ph) { return (osc)floor(.5 + O_SAMP_MAX * cos(ph)); }  /* :128 */
static inline osc o_sin(double ph) { return (osc)floor(.5 + O_SAMP_MAX * sin(ph)); }  /* :129 */
#else
static inline osc o_smul(osc a, osc b) { return a * b; }                            /* :86 */
static inline osc o_half(osc x) { return x * (osc).5; }                             /* :138 */
static inline osc o_half_sum(osc a, osc b) { return (a + b) * (osc).5; }
static inline osc o_half_dif(osc a, osc b) { return (a - b) * (osc).5; }
static inline ocpx o_cmul(ocpx a, ocpx b)                                           /* :87-89 */
{
    ocpx m;
    m.r = a.r * b.r - a.i * b.i;
    m.i = a.r * b.i + a.i * b.r;
    return m;
}
static inline ocpx o_fixdiv(ocpx c, int k) { (void)k; return c; }                   /* :90 no-op */
static inline osc o_cos(double ph) { return (osc)cos(ph); }                         /* :136 */
static inline osc o_sin(double ph) { return (osc)sin(ph); }                         /* :137 */
#endif

/* C_ADD / C_SUB (_kiss_fft_guts.h:100-112): plain ops stored back to the scalar type */
static inline ocpx o_add(ocpx a, ocpx b) { ocpx c; c.r = (osc)(a.r + b.r); c.i = (osc)(a.i + b.i); return c; }
static inline ocpx o_sub(ocpx a, ocpx b) { ocpx c; c.r = (osc)(a.r - b.r); c.i = (osc)(a.i - b.i); return c; }

/* ---- plan ---------------------------------------------------------------------------------- */
typedef struct {
    int nfft, inverse, nstages;
    int p[O_MAXSTAGES], m[O_MAXSTAGES];
    ocpx *tw; /* nfft entries */
} oplan;

/* radix schedule: 4s, then 2s, then odd numbers upward, cut at floor(sqrt(n)) (kiss_fft.c:306-328).
 * Returns the number of stages; facbuf receives p0,m0,p1,m1,... like the reference's factors[]. */
int oracle_factor(int n, int *facbuf)
{
    int p = 4, ns = 0;
    double root = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if (p > root) p = n;
        }
        n /= p;
        facbuf[2 * ns] = p;
        facbuf[2 * ns + 1] = n;
        ++ns;
    } while (n > 1);
    return ns;
}

/* twiddle table tw[i] = exp(-+ 2 pi j i / nfft) from double libm (kiss_fft.c:361-367) */
void oracle_twiddles(int nfft, int inverse, ocpx *tw)
{
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < nfft; ++i) {
        double phase = -2 * pi * i / nfft;
        if (inverse) phase *= -1;
        tw[i].r = o_cos(phase);
        tw[i].i = o_sin(phase);
    }
}

/* real-FFT split twiddles, ncfft/2 of them (kiss_fftr.c:53-59) */
void oracle_super_twiddles(int ncfft, int inverse, ocpx *st)
{
    for (int i = 0; i < ncfft / 2; ++i) {
        double phase = -3.14159265358979323846264338327 * ((double)(i + 1) / ncfft + .5);
        if (inverse) phase *= -1;
        st[i].r = o_cos(phase);
        st[i].i = o_sin(phase);
    }
}

static oplan *oplan_new(int nfft, int inverse)
{
    oplan *pl = (oplan *)malloc(sizeof(oplan));
    int fac[2 * O_MAXSTAGES];
    pl->nfft = nfft;
    pl->inverse = inverse;
    pl->nstages = oracle_factor(nfft, fac);
    for (int s = 0; s < pl->nstages; ++s) { pl->p[s] = fac[2 * s]; pl->m[s] = fac[2 * s + 1]; }
    pl->tw = (ocpx *)malloc(sizeof(ocpx) * (size_t)nfft);
    oracle_twiddles(nfft, inverse, pl->tw);
    return pl;
}
static void oplan_free(oplan *pl) { free(pl->tw); free(pl); }

/* ---- butterflies on p gathered operands v[0..p-1]; k is the position inside the m-block,
 * fs the twiddle stride F_s; results overwrite v[r] = output k + r*m ----------------------- */

/* kiss_fft.c:15-36 */
static void o_bfly2(ocpx *v, const ocpx *tw, size_t k, size_t fs)
{
    ocpx a = o_fixdiv(v[0], 2), b = o_fixdiv(v[1], 2);
    ocpx t = o_cmul(b, tw[k * fs]);
    v[1] = o_sub(a, t);
    v[0] = o_add(a, t);
}

/* kiss_fft.c:38-84 */
static void o_bfly4(ocpx *v, const ocpx *tw, size_t k, size_t fs, int inverse)
{
    ocpx f0 = o_fixdiv(v[0], 4), f1 = o_fixdiv(v[1], 4), f2 = o_fixdiv(v[2], 4), f3 = o_fixdiv(v[3], 4);
    ocpx s0 = o_cmul(f1, tw[k * fs]);
    ocpx s1 = o_cmul(f2, tw[2 * k * fs]);
    ocpx s2 = o_cmul(f3, tw[3 * k * fs]);
    ocpx s5 = o_sub(f0, s1);
    f0 = o_add(f0, s1);
    ocpx s3 = o_add(s0, s2);
    ocpx s4 = o_sub(s0, s2);
    v[2] = o_sub(f0, s3);
    v[0] = o_add(f0, s3);
    if (inverse) {
        v[1].r = (osc)(s5.r - s4.i); v[1].i = (osc)(s5.i + s4.r);
        v[3].r = (osc)(s5.r + s4.i); v[3].i = (osc)(s5.i - s4.r);
    } else {
        v[1].r = (osc)(s5.r + s4.i); v[1].i = (osc)(s5.i - s4.r);
        v[3].r = (osc)(s5.r - s4.i); v[3].i = (osc)(s5.i + s4.r);
    }
}

/* kiss_fft.c:86-128 */
static void o_bfly3(ocpx *v, const ocpx *tw, size_t k, size_t fs, size_t m)
{
    ocpx epi3 = tw[fs * m];
    ocpx f0 = o_fixdiv(v[0], 3), f1 = o_fixdiv(v[1], 3), f2 = o_fixdiv(v[2], 3);
    ocpx s1 = o_cmul(f1, tw[k * fs]);
    ocpx s2 = o_cmul(f2, tw[2 * k * fs]);
    ocpx s3 = o_add(s1, s2);
    ocpx s0 = o_sub(s1, s2);
    ocpx fm;
    fm.r = (osc)(f0.r - o_half(s3.r));
    fm.i = (osc)(f0.i - o_half(s3.i));
    s0.r = o_smul(s0.r, epi3.i);   /* C_MULBYSCALAR */
    s0.i = o_smul(s0.i, epi3.i);
    v[0] = o_add(f0, s3);
    v[2].r = (osc)(fm.r + s0.i); v[2].i = (osc)(fm.i - s0.r);
    v[1].r = (osc)(fm.r - s0.i); v[1].i = (osc)(fm.i + s0.r);
}

/* kiss_fft.c:130-189 */
static void o_bfly5(ocpx *v, const ocpx *tw, size_t k, size_t fs, size_t m)
{
    ocpx ya = tw[fs * m], yb = tw[fs * 2 * m];
    ocpx f0 = o_fixdiv(v[0], 5), f1 = o_fixdiv(v[1], 5), f2 = o_fixdiv(v[2], 5), f3 = o_fixdiv(v[3], 5),
         f4 = o_fixdiv(v[4], 5);
    ocpx s1 = o_cmul(f1, tw[k * fs]);
    ocpx s2 = o_cmul(f2, tw[2 * k * fs]);
    ocpx s3 = o_cmul(f3, tw[3 * k * fs]);
    ocpx s4 = o_cmul(f4, tw[4 * k * fs]);
    ocpx s7 = o_add(s1, s4), s10 = o_sub(s1, s4), s8 = o_add(s2, s3), s9 = o_sub(s2, s3);
    ocpx s5, s6, s11, s12;

    v[0].r = (osc)(f0.r + (s7.r + s8.r));
    v[0].i = (osc)(f0.i + (s7.i + s8.i));

    s5.r = (osc)(f0.r + o_smul(s7.r, ya.r) + o_smul(s8.r, yb.r));
    s5.i = (osc)(f0.i + o_smul(s7.i, ya.r) + o_smul(s8.i, yb.r));
    s6.r = (osc)(o_smul(s10.i, ya.i) + o_smul(s9.i, yb.i));
    s6.i = (osc)(-o_smul(s10.r, ya.i) - o_smul(s9.r, yb.i));
    v[1] = o_sub(s5, s6);
    v[4] = o_add(s5, s6);

    s11.r = (osc)(f0.r + o_smul(s7.r, yb.r) + o_smul(s8.r, ya.r));
    s11.i = (osc)(f0.i + o_smul(s7.i, yb.r) + o_smul(s8.i, ya.r));
    s12.r = (osc)(-o_smul(s10.i, yb.i) + o_smul(s9.i, ya.i));
    s12.i = (osc)(o_smul(s10.r, yb.i) - o_smul(s9.r, ya.i));
    v[2] = o_add(s11, s12);
    v[3] = o_sub(s11, s12);
}

/* kiss_fft.c:192-233: O(p^2) DFT, twiddle index walked modulo nfft, accumulation order q=1..p-1 */
static void o_bfly_generic(ocpx *v, const ocpx *tw, size_t u, size_t fs, size_t m, int p, int nfft)
{
    ocpx sc[O_MAXSTAGES * 64];
    ocpx *scratch = sc;
    ocpx *heap = NULL;
    if ((size_t)p > sizeof(sc) / sizeof(sc[0])) scratch = heap = (ocpx *)malloc(sizeof(ocpx) * (size_t)p);
    for (int q1 = 0; q1 < p; ++q1) scratch[q1] = o_fixdiv(v[q1], p);
    size_t k = u;
    for (int q1 = 0; q1 < p; ++q1) {
        size_t twidx = 0;
        ocpx acc = scratch[0];
        for (int q = 1; q < p; ++q) {
            twidx += fs * k;
            if (twidx >= (size_t)nfft) twidx -= (size_t)nfft;
            acc = o_add(acc, o_cmul(scratch[q], tw[twidx]));
        }
        v[q1] = acc;
        k += m;
    }
    free(heap);
}

/* One transform: breadth-first passes, innermost stage (last factor) first == the order in which
 * the reference's recursion actually executes its butterflies (kiss_fft.c:235-300). */
static void oplan_exec(const oplan *pl, const ocpx *fin, ocpx *fout, size_t in_stride, ocpx *w0, ocpx *w1)
{
    const int N = pl->nfft, L = pl->nstages;
    ocpx *cur = w0, *nxt = w1;
    ocpx stackv[64];
    ocpx *v = stackv, *heap = NULL;
    /* level L: Y_L[off][0] = x[off]  (the leaf copies, kiss_fft.c:274-278) */
    for (int i = 0; i < N; ++i) cur[i] = fin[(size_t)i * in_stride];
    size_t F = (size_t)N; /* F_L */
    for (int s = L - 1; s >= 0; --s) {
        const int p = pl->p[s];
        const size_t m = (size_t)pl->m[s];
        F /= (size_t)p; /* F_s */
        if (p > 64 && !heap) v = heap = (ocpx *)malloc(sizeof(ocpx) * (size_t)N);
        for (size_t k = 0; k < m; ++k)
            for (size_t off = 0; off < F; ++off) {
                for (int q = 0; q < p; ++q) v[q] = cur[(k * p + q) * F + off];
                switch (p) {
                case 2: o_bfly2(v, pl->tw, k, F); break;
                case 3: o_bfly3(v, pl->tw, k, F, m); break;
                case 4: o_bfly4(v, pl->tw, k, F, pl->inverse); break;
                case 5: o_bfly5(v, pl->tw, k, F, m); break;
                default: o_bfly_generic(v, pl->tw, k, F, m, p, N); break;
                }
                for (int r = 0; r < p; ++r) nxt[(k + r * m) * F + off] = v[r];
            }
        ocpx *t = cur; cur = nxt; nxt = t;
    }
    memcpy(fout, cur, sizeof(ocpx) * (size_t)N);
    free(heap);
}

/* ---- exported entry points ------------------------------------------------------------------ */

/* kiss_fft_stride (kiss_fft.c:375-399), `howmany` transforms, input element stride in_stride,
 * consecutive transforms in_dist / out_dist elements apart.  fin may equal fout. */
void oracle_fft_batch(int nfft, int inverse, const ocpx *fin, ocpx *fout, int in_stride, size_t howmany,
                      size_t in_dist, size_t out_dist)
{
    oplan *pl = oplan_new(nfft, inverse);
    ocpx *w0 = (ocpx *)malloc(sizeof(ocpx) * (size_t)nfft * 2);
    for (size_t b = 0; b < howmany; ++b)
        oplan_exec(pl, fin + b * in_dist, fout + b * out_dist, (size_t)in_stride, w0, w0 + nfft);
    free(w0);
    oplan_free(pl);
}

void oracle_fft(int nfft, int inverse, const ocpx *fin, ocpx *fout, int in_stride)
{
    oracle_fft_batch(nfft, inverse, fin, fout, in_stride, 1, 0, 0);
}

/* kiss_fftr (kiss_fftr.c:63-117): nfft real -> nfft/2+1 complex; rows are nfft scalars /
 * nfft/2+1 complex apart. Returns -1 if nfft is odd (kiss_fftr.c:29-32). */
int oracle_fftr_batch(int nfft, const osc *timedata, ocpx *freqdata, size_t howmany)
{
    if (nfft & 1) return -1;
    const int nc = nfft / 2;
    oplan *pl = oplan_new(nc, 0);
    ocpx *st = (ocpx *)malloc(sizeof(ocpx) * (size_t)(nc / 2 + 1));
    ocpx *T = (ocpx *)malloc(sizeof(ocpx) * (size_t)nc * 3);
    oracle_super_twiddles(nc, 0, st);
    for (size_t b = 0; b < howmany; ++b) {
        ocpx *out = freqdata + b * (size_t)(nc + 1);
        oplan_exec(pl, (const ocpx *)(timedata + b * (size_t)nfft), T, 1, T + nc, T + 2 * nc);
        ocpx tdc = o_fixdiv(T[0], 2);
        out[0].r = (osc)(tdc.r + tdc.i);
        out[nc].r = (osc)(tdc.r - tdc.i);
        out[0].i = out[nc].i = 0;
        for (int k = 1; k <= nc / 2; ++k) {
            ocpx fpk = T[k], fpnk;
            fpnk.r = T[nc - k].r;
            fpnk.i = (osc)(-T[nc - k].i);
            fpk = o_fixdiv(fpk, 2);
            fpnk = o_fixdiv(fpnk, 2);
            ocpx f1k = o_add(fpk, fpnk), f2k = o_sub(fpk, fpnk);
            ocpx tw = o_cmul(f2k, st[k - 1]);
            out[k].r = o_half_sum(f1k.r, tw.r);
            out[k].i = o_half_sum(f1k.i, tw.i);
            out[nc - k].r = o_half_dif(f1k.r, tw.r);
            out[nc - k].i = o_half_dif(tw.i, f1k.i);
        }
    }
    free(T); free(st); oplan_free(pl);
    return 0;
}

/* kiss_fftri (kiss_fftr.c:119-155): nfft/2+1 complex -> nfft real */
int oracle_fftri_batch(int nfft, const ocpx *freqdata, osc *timedata, size_t howmany)
{
    if (nfft & 1) return -1;
    const int nc = nfft / 2;
    oplan *pl = oplan_new(nc, 1);
    ocpx *st = (ocpx *)malloc(sizeof(ocpx) * (size_t)(nc / 2 + 1));
    ocpx *T = (ocpx *)malloc(sizeof(ocpx) * (size_t)nc * 3);
    oracle_super_twiddles(nc, 1, st);
    for (size_t b = 0; b < howmany; ++b) {
        const ocpx *F = freqdata + b * (size_t)(nc + 1);
        T[0].r = (osc)(F[0].r + F[nc].r);
        T[0].i = (osc)(F[0].r - F[nc].r);
        T[0] = o_fixdiv(T[0], 2);
        for (int k = 1; k <= nc / 2; ++k) {
            ocpx fk = F[k], fnkc;
            fnkc.r = F[nc - k].r;
            fnkc.i = (osc)(-F[nc - k].i);
            fk = o_fixdiv(fk, 2);
            fnkc = o_fixdiv(fnkc, 2);
            ocpx fek = o_add(fk, fnkc), tmp = o_sub(fk, fnkc);
            ocpx fok = o_cmul(tmp, st[k - 1]);
            T[k] = o_add(fek, fok);
            T[nc - k] = o_sub(fek, fok);
            T[nc - k].i = (osc)(-T[nc - k].i);   /* kiss_fftr.c:149-153: "*= -1" */
        }
        oplan_exec(pl, T, (ocpx *)(timedata + b * (size_t)nfft), 1, T + nc, T + 2 * nc);
    }
    free(T); free(st); oplan_free(pl);
    return 0;
}

/* kiss_fftnd (kiss_fftnd.c:156-188): for axis k = 0..ndims-1 view the buffer as dims[k] x stride,
 * transform every column and store it as a row.  fin may equal fout. */
void oracle_fftnd(const int *dims, int ndims, int inverse, const ocpx *fin, ocpx *fout)
{
    size_t total = 1;
    int maxd = 1;
    for (int i = 0; i < ndims; ++i) { total *= (size_t)dims[i]; if (dims[i] > maxd) maxd = dims[i]; }
    ocpx *a = (ocpx *)malloc(sizeof(ocpx) * total);
    ocpx *b = (ocpx *)malloc(sizeof(ocpx) * total);
    ocpx *w = (ocpx *)malloc(sizeof(ocpx) * (size_t)maxd * 2);
    memcpy(a, fin, sizeof(ocpx) * total);
    for (int k = 0; k < ndims; ++k) {
        const size_t n = (size_t)dims[k], stride = total / n;
        oplan *pl = oplan_new(dims[k], inverse);
        for (size_t i = 0; i < stride; ++i) oplan_exec(pl, a + i, b + i * n, stride, w, w + maxd);
        oplan_free(pl);
        ocpx *t = a; a = b; b = t;
    }
    memcpy(fout, a, sizeof(ocpx) * total);
    free(a); free(b); free(w);
}

/* kiss_fftndr (kiss_fftndr.c:86-110): real transform along the LAST axis, bin-major scatter, then the
 * complex N-D transform over the remaining axes for every bin. out is [dimOther][dimReal/2+1]. */
int oracle_fftndr(const int *dims, int ndims, const osc *timedata, ocpx *freqdata)
{
    const int dimReal = dims[ndims - 1];
    if (dimReal & 1) return -1;
    size_t dimOther = 1;
    for (int i = 0; i < ndims - 1; ++i) dimOther *= (size_t)dims[i];
    const size_t nrbins = (size_t)dimReal / 2 + 1;
    ocpx *rows = (ocpx *)malloc(sizeof(ocpx) * dimOther * nrbins);
    ocpx *binmajor = (ocpx *)malloc(sizeof(ocpx) * dimOther * nrbins);
    ocpx *tmp = (ocpx *)malloc(sizeof(ocpx) * dimOther);
    oracle_fftr_batch(dimReal, timedata, rows, dimOther);
    for (size_t k1 = 0; k1 < dimOther; ++k1)
        for (size_t k2 = 0; k2 < nrbins; ++k2) binmajor[k2 * dimOther + k1] = rows[k1 * nrbins + k2];
    for (size_t k2 = 0; k2 < nrbins; ++k2) {
        if (ndims > 1) oracle_fftnd(dims, ndims - 1, 0, binmajor + k2 * dimOther, tmp);
        else memcpy(tmp, binmajor + k2 * dimOther, sizeof(ocpx) * dimOther);
        for (size_t k1 = 0; k1 < dimOther; ++k1) freqdata[k1 * nrbins + k2] = tmp[k1];
    }
    free(rows); free(binmajor); free(tmp);
    return 0;
}

/* kiss_fftndri (kiss_fftndr.c:112-132) */
int oracle_fftndri(const int *dims, int ndims, const ocpx *freqdata, osc *timedata)
{
    const int dimReal = dims[ndims - 1];
    if (dimReal & 1) return -1;
    size_t dimOther = 1;
    for (int i = 0; i < ndims - 1; ++i) dimOther *= (size_t)dims[i];
    const size_t nrbins = (size_t)dimReal / 2 + 1;
    ocpx *binmajor = (ocpx *)malloc(sizeof(ocpx) * dimOther * nrbins);
    ocpx *rows = (ocpx *)malloc(sizeof(ocpx) * dimOther * nrbins);
    ocpx *tmp = (ocpx *)malloc(sizeof(ocpx) * dimOther);
    for (size_t k2 = 0; k2 < nrbins; ++k2) {
        for (size_t k1 = 0; k1 < dimOther; ++k1) tmp[k1] = freqdata[k1 * nrbins + k2];
        if (ndims > 1) oracle_fftnd(dims, ndims - 1, 1, tmp, binmajor + k2 * dimOther);
        else memcpy(binmajor + k2 * dimOther, tmp, sizeof(ocpx) * dimOther);
    }
    for (size_t k1 = 0; k1 < dimOther; ++k1)
        for (size_t k2 = 0; k2 < nrbins; ++k2) rows[k1 * nrbins + k2] = binmajor[k2 * dimOther + k1];
    oracle_fftri_batch(dimReal, rows, timedata, dimOther);
    free(binmajor); free(rows); free(tmp);
    return 0;
}

/* kiss_fft_next_fast_size (kiss_fft.c:412-424) */
int oracle_next_fast_size(int n)
{
    for (;; ++n) {
        int m = n;
        while ((m % 2) == 0) m /= 2;
        while ((m % 3) == 0) m /= 3;
        while ((m % 5) == 0) m /= 5;
        if (m <= 1) return n;
    }
}

int oracle_sizeof_scalar(void) { return (int)sizeof(osc); }
