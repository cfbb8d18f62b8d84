// kf_body.h -- bodies of the batched transform kernels, written against an execution-environment concept.
//
//   fused_body<A, PT, MODE>   compile-time plan PT::D (kf_plan.h): every transform is read from HBM once and
//                             written once; the radix stages of kf_work/kf_bfly* (kiss_fft.c:15-300) run as
//                             register groups that exchange through shared memory.  Persistent CTAs
//                             (grid = resident CTAs), `tpc` transforms per CTA iteration.
//   generic_body<A>           run-time plan for any nfft whose two exchange buffers fit in shared memory:
//                             one radix stage per pass over shared memory.
//
// MODE selects what surrounds the complex transform:
//   kC2C      kiss_fft / kiss_fft_stride with unit or small input stride             (kiss_fft.c:375-404)
//   kC2CCol   kiss_fftnd axis pass: `tpc` adjacent columns are loaded together so that HBM reads are
//             tpc*sizeof(cpx)-byte segments, each column is written as a contiguous row (kiss_fftnd.c:172-178)
//   kC2CColTw / kC2CColCol   the two steps of the four-step transform of a long row N = N1*N2 (float / double): columns
//             in, rows out multiplied by W_N^(column * k); columns in, columns out (natural order).  kC2CColCol alone is
//             also an axis pass of an N-D array that leaves the layout unchanged (every datatype)
//   kR2C      kiss_fftr: packed complex transform + split-twiddle post pass fused     (kiss_fftr.c:63-117)
//   kC2R      kiss_fftri: split pre pass fused + inverse complex transform            (kiss_fftr.c:119-155)
#pragma once
#include <stddef.h>

#include "kf_engine.h"

namespace kf {

enum Mode { kC2C = 0, kC2CCol = 1, kR2C = 2, kC2R = 3, kC2CColTw = 4, kC2CColCol = 5 };
constexpr bool is_col_mode(int m) { return m == kC2CCol || m == kC2CColTw || m == kC2CColCol; }

template <class A>
struct KParams {
    // column modes with a tensor-map input ring (PlanDesc::nstage > 0): the CUtensorMap of the input array viewed as
    // [planes][nfft rows][columns] (opaque here; encoded by the launcher, read by cp.async.bulk.tensor).  First member
    // so that it sits on the 64-byte boundary the hardware wants.
    alignas(64) unsigned char tmap[128];
    const typename A::C* in;     // complex view of the input (kR2C: the real rows, 2 scalars per element)
    typename A::C* out;          // complex view of the output (kC2R: the real rows)
    long long howmany;
    long long in_dist, out_dist; // distance between consecutive transforms, in complex elements of each side
    long long in_stride;         // element stride of the input (kiss_fft_stride's in_stride)
    // column mode only: transforms are numbered plane-major, `ncols` columns per plane; plane p starts at
    // in + p*in_pdist / out + p*out_pdist and column c of it at + c*in_dist / + c*out_dist.  ncols == 0: one plane.
    long long ncols, in_pdist, out_pdist;
    // column mode, fused exchange: when npeers > 0 the columns of a plane are split into npeers blocks of `cols_per_peer`
    // and block s is written through peer[s] (a mapped pointer into rank s's receive buffer: NVLink peer stores)
    // peer_col_dist: input columns between the blocks of consecutive peers (== cols_per_peer when the blocks are adjacent;
    // larger when only a chunk of every peer's column range is processed per launch)
    typename A::C* peer[16];
    long long cols_per_peer, peer_col_dist;
    int npeers;
    int max_ctas;                // host side only: cap on the persistent grid of this launch (0 = fill the device)
    KF_HD long long in_col(long long col) const   // input column index of column `col` of a plane
    {
        return (npeers > 0 && peer_col_dist != cols_per_peer) ? (col / cols_per_peer) * peer_col_dist + col % cols_per_peer : col;
    }
    KF_HD long long in_off(long long b) const { return ncols > 0 ? (b / ncols) * in_pdist + in_col(b % ncols) * in_dist : b * in_dist; }
    KF_HD long long out_off(long long b) const { return ncols > 0 ? (b / ncols) * out_pdist + (b % ncols) * out_dist : b * out_dist; }
    KF_HD typename A::C* out_ptr(long long b) const
    {
        if (npeers <= 0) return out + out_off(b);
        const long long pl = b / ncols, col = b % ncols;
        return peer[col / cols_per_peer] + pl * out_pdist + (col % cols_per_peer) * out_dist;
    }
    const typename A::C* tw;     // N twiddles (kR2C/kC2R: of the ncfft-point sub-transform)
    const typename A::C* stw;    // ncfft/2 split twiddles (kiss_fftr.c:53-59), real modes only
    const typename A::C* gtw;    // per-group stage-twiddle tables of the fused plan (kf_twtab.h), unused by the generic kernel
    cx<typename A::R> g0tw[kMaxG0Slots];   // group 0's stage twiddles (plan constants -> constant bank)
    cx<typename A::R> ctw[kMaxCtw];        // split-twiddle constants (PlanDesc::twmode == 1)
    PlanConsts<A> pc;
    int inverse;
};

template <class C>
KF_HD C ld_stream(const C* p)
{
#if !defined(__CUDA_ARCH__)
    return *p;
#else
    // input rows are read exactly once: bypass L1 allocation.  Deliberately NOT the non-coherent (.nc) path: the API
    // allows in-place operation (fin == fout, kiss_fft.c:377-395; every in-layout axis pass), i.e. the same kernel writes
    // this memory, which PTX leaves undefined for .nc loads.  In-place safety rests on ownership: a tile is read only by
    // the CTA that later writes it, and read completely (into registers / shared memory) before its first store.
    if constexpr (sizeof(C) == 4) {
        unsigned u;
        asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(u) : "l"(p));
        return *reinterpret_cast<C*>(&u);
    } else if constexpr (sizeof(C) == 8) {
        unsigned long long u;
        asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(u) : "l"(p));
        return *reinterpret_cast<C*>(&u);
    } else {
        unsigned long long u0, u1;
        asm volatile("ld.global.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(u0), "=l"(u1) : "l"(p));
        struct alignas(16) U2 { unsigned long long a, b; } u{u0, u1};
        return *reinterpret_cast<C*>(&u);
    }
#endif
}

template <class A, bool UNIT>
struct SrcGlobal {
    const typename A::C* base;
    long long stride;   // ignored when UNIT (contiguous rows: element offsets fold into the load immediates)
    KF_HD cx<typename A::R> load(int i) const
    {
        if constexpr (UNIT) return A::load(ld_stream(base + i));
        else return A::load(ld_stream(base + (long long)i * stride));
    }
    template <int IT_, int E_>
    KF_HD cx<typename A::R> get(int i) const { return load(i); }
};
template <class A>
struct SrcShared {
    const typename A::C* base;   // natural order, unpadded
    KF_HD cx<typename A::R> load(int i) const { return A::load(base[i]); }
    template <int IT_, int E_>
    KF_HD cx<typename A::R> get(int i) const { return load(i); }
};
// column modes with the tensor-map ring: the tile landed as [nfft rows][ncol adjacent columns]; element i of this
// thread's column is `ncol` elements further down
template <class A>
struct SrcStageCol {
    const typename A::C* base;   // stage + column index
    int ncol;
    KF_HD cx<typename A::R> load(int i) const { return A::load(base[i * ncol]); }
    template <int IT_, int E_>
    KF_HD cx<typename A::R> get(int i) const { return load(i); }
};
// kiss_fftri's split pre pass (kiss_fftr.c:131-153) fused into the first group's loads: element k of the packed
// complex input T[] of the inverse transform is computed on the fly from the half spectrum F[k], F[nc-k] and the
// split twiddle.  Every T[k] is produced independently (the reference produces T[k] and T[nc-k] together; here
// the thread that needs T[nc-k] recomputes the pair) -- same operands, same roundings, no T[] round trip.
template <class A, bool SHARED>
struct SrcC2RFused {
    const typename A::C* row;   // F[0..nc]: landed stage (SHARED) or global memory
    const typename A::C* stw;
    int nc;
    KF_HD cx<typename A::R> f(int i) const
    {
        if constexpr (SHARED) return A::load(row[i]);
        else return A::load(ld_stream(row + i));
    }
    KF_HD cx<typename A::R> load(int k) const
    {
        typedef cx<typename A::R> X;
        const bool lo = 2 * k < nc;          // k == nc/2 takes the second assignment, which wins in the reference
        const int ks = lo ? k : nc - k;
        const X Fa = f(ks), Fb = f(nc - ks);
        const X st = (ks == 0) ? Fa : A::load(TwTab<A>::ro_load_c(stw + (ks - 1)));
        X Tk, Tnk;
        fftri_pre_pair<A>(ks, Fa, Fb, st, Tk, Tnk);
        return lo ? Tk : Tnk;
    }
    template <int IT_, int E_>
    KF_HD cx<typename A::R> get(int k) const { return load(k); }
};

// Destinations of the last group: put<it, e>(k, v) receives output element k, produced from register e of the
// thread's it-th work item.
template <class A>
struct DstGlobal {
    typename A::C* base;
    template <int IT_, int E_>
    KF_HD void put(int k, const cx<typename A::R>& v) const { base[k] = A::store(v); }
};
// four-step, first step: row `col` of the intermediate array, multiplied by the long transform's twiddle W_N^(col*k)
template <class A>
struct DstGlobalTw {
    typename A::C* base;
    const typename A::C* twbig;
    long long col;
    template <int IT_, int E_>
    KF_HD void put(int k, const cx<typename A::R>& v) const
    {
        base[k] = A::store(A::cmul(v, A::load(TwTab<A>::ro_load_c(twbig + col * k))));
    }
};
// four-step, second step: element k of a column goes `stride` elements down (adjacent lanes hold adjacent columns)
template <class A>
struct DstGlobalStrided {
    typename A::C* base;
    long long stride;
    template <int IT_, int E_>
    KF_HD void put(int k, const cx<typename A::R>& v) const { base[k * stride] = A::store(v); }
};
template <class A>
struct DstShared {
    typename A::C* base;         // natural order, unpadded
    template <int IT_, int E_>
    KF_HD void put(int k, const cx<typename A::R>& v) const { base[k] = A::store(v); }
};
// keeps the outputs in the caller's registers (kiss_fftr post pass by warp shuffles)
template <class A, int R_>
struct DstRegs {
    cx<typename A::R>* regs;
    template <int IT_, int E_>
    KF_HD void put(int, const cx<typename A::R>& v) const { regs[IT_ * R_ + E_] = v; }
};
struct NoSrc {
    template <class T = int>
    KF_HD cx<int> load(int) const { return cx<int>{0, 0}; }
};

// ---- group sequencing: group g reads exchange buffer (x0+g-1)&1 and writes (x0+g)&1; one barrier per exchange ----
template <class A, PlanDesc D, int g, class Src, class Dst, class Env, int GEND = D.G>
KF_HD void run_groups(Env& env, int t, bool active, const Src& src, const Dst& dst, typename A::C* buf0,
                      typename A::C* buf1, const TwTab<A>& tw, const PlanConsts<A>& pc, int inverse)
{
    // buf0 is read by this group (unused for g == 0), buf1 is written (unused for the last group); groups g..GEND-1
    if constexpr (g < GEND) {
        run_group<A, D, g, Src, Dst>(t, active, src, dst, buf0, buf1, tw, pc, inverse);
        if constexpr (g + 1 < D.G) env.sync();
        if constexpr (g + 1 < GEND) run_groups<A, D, g + 1, Src, Dst, Env, GEND>(env, t, active, src, dst, buf1, buf0, tw, pc, inverse);
    }
}

// ---------------------------------------------------------------------------------------------------------
// kiss_fftr: last group + split post pass (kiss_fftr.c:88-116) on PAIRS of work items.
// The last group has m = nc/R work items; item u produces T[u + j*m], j < R.  The bin pair of T[u + j*m] is
// T[nc - u - j*m] = T[(m-u) + (R-1-j)*m], an output of item m-u.  A thread therefore runs items u and m-u together
// and finds every pair (k, nc-k) of the post pass complete in its registers: no T[] round trip through shared
// memory.  Items 0 and m/2 pair with themselves and are taken together by the thread whose u is 0.
// Same operands and roundings as the reference's loop over k = 1..nc/2, so fixed point stays bit-exact.
// ---------------------------------------------------------------------------------------------------------
// Bin pairs a thread of a paired group handles, in the order both paired passes enumerate them: pair jj (< R) of the
// thread whose first item is u, with m work items in the group (half = m/2, nc = R*m bins):
//   u != 0:  jj even -> k = u + (jj/2)*m        jj odd -> k = (m-u) + (jj/2)*m
//   u == 0:  jj == 0 -> k = nc/2 (self-paired)  jj even -> k = (jj/2)*m      jj odd -> k = half + (jj/2)*m
KF_HD int pair_ks(int u, int jj, int m, int nc)
{
    const int e = jj >> 1, odd = jj & 1;
    if (u != 0) return (odd ? m - u : u) + e * m;
    return jj == 0 ? nc / 2 : (odd ? m / 2 : 0) + e * m;
}

// split twiddle of pair jj: from the thread's registers when hoisted (HS), else from the table
template <class A, bool HS>
KF_HD cx<typename A::R> split_tw(const typename A::C* stw, const cx<typename A::R>* hst, int idx, int ks)
{
    if constexpr (HS) return hst[idx];
    else return A::load(TwTab<A>::ro_load_c(stw + (ks - 1)));
}

template <class A>
KF_HD void r2c_emit_pair(int ks, int nc, const cx<typename A::R>& Tk, const cx<typename A::R>& Tnk, const cx<typename A::R>& st,
                         typename A::C* out)
{
    typedef cx<typename A::R> X;
    X ok, onk;
    fftr_post_pair<A>(ks, nc, Tk, Tnk, st, ok, onk);
    if (ks != nc - ks) out[ks] = A::store(ok);     // ks == nc/2: the reference's second assignment wins
    out[nc - ks] = A::store(onk);
}

// `mid()` runs between the loads of the exchange buffer and the butterflies (a no-op, or the barrier + refill of the
// in-place input stage; it contains a CTA barrier only when every thread reaches it exactly once, kIt == 1).
template <class A, PlanDesc D, class Mid>
KF_HD void run_r2c_last_paired(int t, bool active, const typename A::C* rd, const TwTab<A>& tw, const PlanConsts<A>& pc,
                               const typename A::C* stw, const cx<typename A::R>* hst, typename A::C* out, int inverse, const Mid& mid)
{
    typedef cx<typename A::R> X;
    constexpr int gl = D.G - 1, R = D.R(gl), m = D.items(gl), nc = D.N, half = m / 2;
    constexpr int kIt = (half + D.team - 1) / D.team;                   // u = 0 .. half-1
    constexpr bool HS = (D.hoist & 1) != 0;
    const typename A::R sg = A::sign_of(inverse);
    static_for<kIt>([&](auto ITER) {
        constexpr int it = decltype(ITER)::value;
        const int u = t + it * D.team;
        const bool on = active && u < half;
        const int wa = on ? u : 0, wb = (on && u != 0) ? m - u : half;   // items (u, m-u); (0, m/2) for u == 0
        constexpr int kHa = (D.hoist & 4) ? D.hoist_base(gl, 2 * it) : -1, kHb = (D.hoist & 4) ? D.hoist_base(gl, 2 * it + 1) : -1;
        auto st = [&](int jj, int ks) { return split_tw<A, HS>(stw, hst, it * R + jj, ks); };
        X a[R], b[R];
        if (on) {
            item_load<A, D, gl>(wa, rd, a);
            item_load<A, D, gl>(wb, rd, b);
        }
        mid();
        if (on) {
            item_stages<A, D, gl, kHa>(wa, a, tw, pc, sg);
            item_stages<A, D, gl, kHb>(wb, b, tw, pc, sg);
            if (u == 0) {
                // a = T[j*m], b = T[m/2 + j*m]: both items pair with themselves
                X ok, onk;
                constexpr int e0 = D.reg_of_j(gl, 0), eh = D.reg_of_j(gl, R / 2);
                fftr_post_pair<A>(0, nc, a[e0], a[e0], a[e0], ok, onk);        // DC and Nyquist bins
                out[0] = A::store(ok);
                out[nc] = A::store(onk);
                static_for<R / 2>([&](auto JJ) {
                    constexpr int j = decltype(JJ)::value;
                    if constexpr (j >= 1) r2c_emit_pair<A>(j * m, nc, a[D.reg_of_j(gl, j)], a[D.reg_of_j(gl, R - j)], st(2 * j, j * m), out);
                    r2c_emit_pair<A>(half + j * m, nc, b[D.reg_of_j(gl, j)], b[D.reg_of_j(gl, R - 1 - j)], st(2 * j + 1, half + j * m), out);
                });
                r2c_emit_pair<A>(nc / 2, nc, a[eh], a[eh], st(0, nc / 2), out);   // k == nc/2 pairs with itself
            } else {
                static_for<R / 2>([&](auto JJ) {
                    constexpr int j = decltype(JJ)::value;
                    constexpr int ej = D.reg_of_j(gl, j), en = D.reg_of_j(gl, R - 1 - j);
                    r2c_emit_pair<A>(u + j * m, nc, a[ej], b[en], st(2 * j, u + j * m), out);
                    r2c_emit_pair<A>((m - u) + j * m, nc, b[ej], a[en], st(2 * j + 1, (m - u) + j * m), out);
                });
            }
        }
    });
}

// ---------------------------------------------------------------------------------------------------------
// kiss_fftri: split pre pass (kiss_fftr.c:131-153) + first group on PAIRS of work items -- the mirror image.
// The first group has W = nc/R work items; item u consumes T[u + e*W], e < R, and T[nc - u - e*W] is input R-1-e of
// item W-u.  The thread that runs items u and W-u reads every F[k], F[nc-k] and split twiddle exactly once and
// produces both T values of the pair as the reference does (SrcC2RFused evaluates each pair twice).
// ---------------------------------------------------------------------------------------------------------
template <class A, class F>
KF_HD void c2r_make_pair(int ks, int nc, const F& f, const cx<typename A::R>& st, cx<typename A::R>& Tk, cx<typename A::R>& Tnk)
{
    typedef cx<typename A::R> X;
    const X Fa = f(ks), Fb = f(nc - ks);
    fftri_pre_pair<A>(ks, Fa, Fb, st, Tk, Tnk);
}

template <class A, PlanDesc D, class F, class Mid>
KF_HD void run_c2r_first_paired(int t, bool active, const F& f, typename A::C* wr, const TwTab<A>& tw, const PlanConsts<A>& pc,
                                const typename A::C* stw, const cx<typename A::R>* hst, int inverse, const Mid& mid)
{
    typedef cx<typename A::R> X;
    constexpr int R = D.R(0), W = D.items(0), nc = D.N, half = W / 2;
    constexpr int kIt = (half + D.team - 1) / D.team;
    constexpr bool HS = (D.hoist & 1) != 0;
    const typename A::R sg = A::sign_of(inverse);
    static_for<kIt>([&](auto ITER) {
        constexpr int it = decltype(ITER)::value;
        const int u = t + it * D.team;
        const bool on = active && u < half;
        const int wa = on ? u : 0, wb = (on && u != 0) ? W - u : half;
        auto st = [&](int jj, int ks) { return split_tw<A, HS>(stw, hst, it * R + jj, ks); };
        X a[R], b[R];
        if (on) {
            if (u == 0) {
                X dummy;
                fftri_pre_pair<A>(0, f(0), f(nc), f(0), a[0], dummy);            // T[0] from F[0], F[nc]
                static_for<R / 2>([&](auto EE) {
                    constexpr int e = decltype(EE)::value;
                    if constexpr (e >= 1) c2r_make_pair<A>(e * W, nc, f, st(2 * e, e * W), a[e], a[R - e]);
                    c2r_make_pair<A>(half + e * W, nc, f, st(2 * e + 1, half + e * W), b[e], b[R - 1 - e]);
                });
                c2r_make_pair<A>(nc / 2, nc, f, st(0, nc / 2), dummy, a[R / 2]);   // k == nc/2: the second assignment wins
            } else {
                static_for<R / 2>([&](auto EE) {
                    constexpr int e = decltype(EE)::value;
                    c2r_make_pair<A>(u + e * W, nc, f, st(2 * e, u + e * W), a[e], b[R - 1 - e]);
                    c2r_make_pair<A>((W - u) + e * W, nc, f, st(2 * e + 1, (W - u) + e * W), b[e], a[R - 1 - e]);
                });
            }
        }
        mid();
        if (on) {
            item_stages<A, D, 0>(wa, a, tw, pc, sg);
            item_stages<A, D, 0>(wb, b, tw, pc, sg);
            item_store<A, D, 0>(wa, wr, a);
            item_store<A, D, 0>(wb, wr, b);
        }
    });
}

// ---------------------------------------------------------------------------------------------------------
// Loop-invariant tables into registers (PlanDesc::hoist).  A persistent CTA's thread t runs the same work items of
// every group for every tile, so the twiddles those items need never change: they are fetched once, before the tile
// loop, instead of once per transform (ncu r01: 690-780 MB of L1 table traffic per launch, more than the payload).
// PAIRG = the group this mode runs on work-item pairs (kiss_fftr: the last one), or -1.
// ---------------------------------------------------------------------------------------------------------
template <class A, PlanDesc D, int PAIRG, int g = 1>
KF_HD void hoist_fill(int t, const TwTab<A>& tw, cx<typename A::R>* hreg)
{
    if constexpr (g < D.G) {
        if constexpr (g == PAIRG) {
            if constexpr ((D.hoist & 4) != 0) {
                constexpr int m = D.items(g), half = m / 2, kIt = (half + D.team - 1) / D.team;
                static_for<kIt>([&](auto ITER) {
                    constexpr int it = decltype(ITER)::value;
                    const int u = t + it * D.team;
                    const bool in = u < half;
                    hoist_item<A, D, g, D.hoist_base(g, 2 * it)>(tw, in ? u : 0, hreg);
                    hoist_item<A, D, g, D.hoist_base(g, 2 * it + 1)>(tw, (in && u != 0) ? m - u : half, hreg);
                });
            }
        } else if constexpr ((D.hoist & 2) != 0) {
            static_for<D.iters(g)>([&](auto ITER) {
                constexpr int it = decltype(ITER)::value;
                const int w = t + it * D.team;
                hoist_item<A, D, g, D.hoist_base(g, it)>(tw, w < D.items(g) ? w : D.items(g) - 1, hreg);
            });
        }
        hoist_fill<A, D, PAIRG, g + 1>(t, tw, hreg);
    }
}

// split twiddles of the bin pairs thread t handles in the paired group with m items of R points (pair_ks order)
template <class A, PlanDesc D, int R, int m>
KF_HD void hoist_fill_split(int t, const typename A::C* stw, cx<typename A::R>* hst)
{
    constexpr int half = m / 2, kIt = (half + D.team - 1) / D.team;
    static_for<kIt>([&](auto ITER) {
        constexpr int it = decltype(ITER)::value;
        const int u0 = t + it * D.team, u = u0 < half ? u0 : 0;
        static_for<R>([&](auto JJ) {
            constexpr int jj = decltype(JJ)::value;
            hst[it * R + jj] = A::load(TwTab<A>::ro_load_c(stw + (pair_ks(u, jj, m, D.N) - 1)));
        });
    });
}

// Shared-memory layout of one CTA of the fused kernel.
//   [exchange buffer A][exchange buffer B]   tpc * pitch elements each (skewed autosort arrays between groups)
//   [input ring: nstage stages]              each stage = one tile of tpc contiguous input rows, natural order,
//                                            filled by a bulk asynchronous copy (TMA, cp.async.bulk) that
//                                            completes on the stage's mbarrier
//   [nstage mbarriers]
template <class A, class PT, int MODE>
struct FusedLayout {
    static constexpr PlanDesc D = PT::D;
    static constexpr bool kRing = D.nstage > 0 && !is_col_mode(MODE);
    // column modes: tiles of tpc adjacent columns x N rows fetched by tensor-map TMA (cp.async.bulk.tensor) into a ring,
    // so the strided HBM reads of the next tiles run under the butterflies of the current one
    static constexpr bool kColRing = D.nstage > 0 && is_col_mode(MODE);
    static constexpr int kBoxRows = D.N < 256 ? D.N : 256;           // tensor-map boxes hold at most 256 rows
    static_assert(!kColRing || D.N % kBoxRows == 0, "column ring: nfft must be a multiple of the 256-row box");
    static constexpr int kRowIn = (MODE == kC2R) ? D.N + 1 : D.N;   // complex elements per input row
    static_assert(D.nbuf == 2 || (D.nbuf == 1 && D.G == 2 && (MODE == kC2C || is_col_mode(MODE))) ||
                      (D.nbuf == 1 && D.G == 3 && D.nstage == 1 && D.paired && (MODE == kR2C || MODE == kC2R)) ||
                      (D.nbuf == 1 && D.G >= 3 && D.nstage == 1 && MODE == kC2C),
                  "single exchange buffer: two-group C2C/column plans, or (one-stage input ring) C2C plans with three or more groups / three-group paired real plans");
    // kiss_fftr post pass by warp shuffles instead of a T[] round trip through shared memory: needs whole warps per
    // team and the last group's work items to divide evenly (PlanDesc::shfl_post asks for it)
    static constexpr bool kShflPost = MODE == kR2C && D.shfl_post && D.team % 32 == 0 && D.items(D.G - 1) % D.team == 0;
    // kiss_fftr with the last group run on work-item pairs (k', m-k'): post pass in registers, no T[] buffer
    static constexpr bool kPairedLast = MODE == kR2C && D.paired && D.G >= 2 && D.R(D.G - 1) % 2 == 0 && D.items(D.G - 1) % 2 == 0;
    // kiss_fftri with the split pre pass + first group run on work-item pairs: every spectrum bin is read once
    static constexpr bool kPairedFirst = MODE == kC2R && D.paired && D.G >= 2 && D.R(0) % 2 == 0 && D.items(0) % 2 == 0;
    // nbuf == 1 on a plan with three or more groups (C2C, or paired real with three): ONE exchange buffer, and the
    // input stage (consumed by the first group) doubles as the second one.  Shared memory per transform in flight drops from 3 rows to 2, so more CTAs
    // are resident; the stage is refilled once the last group has read it (no refill under the butterflies).
    static constexpr bool kStageExch = D.nbuf == 1 && kRing && D.nstage == 1 &&
                                       ((D.G == 3 && ((kPairedLast && D.items(2) / 2 <= D.team) || kPairedFirst)) ||
                                        (D.G >= 3 && MODE == kC2C));
    static constexpr size_t kExchBytes = (D.G >= 2 || (MODE == kR2C && !kShflPost)) ? (size_t)D.nbuf * D.tpc * D.pitch() * sizeof(typename A::C) : 0;
    static constexpr size_t kRingOff = (kExchBytes + 127) / 128 * 128;
    // rows whose byte size is not a multiple of 16 (kiss_fftri: ncfft+1 elements): a tile then starts off the 16-byte
    // grid cp.async.bulk needs, so the copy starts at the aligned address below it and is rounded up to 16 bytes --
    // the tile lands kMis elements into the stage (and up to 15 bytes of the following row come along)
    static constexpr int kE16 = 16 / (int)sizeof(typename A::C) > 0 ? 16 / (int)sizeof(typename A::C) : 1;
    static constexpr bool kSlackRing = kRing && ((size_t)D.tpc * kRowIn * sizeof(typename A::C)) % 16 != 0;
    static constexpr int kLandElems = D.tpc * kRowIn + (kSlackRing ? kE16 : 0);
    static constexpr int kStageElems = (kStageExch && D.tpc * D.pitch() > kLandElems) ? D.tpc * D.pitch() : kLandElems;
    static constexpr size_t kStageBytes = ((size_t)kStageElems * sizeof(typename A::C) + 127) / 128 * 128;
    static constexpr size_t kBarOff = kRingOff + ((kRing || kColRing) ? D.nstage * kStageBytes : 0);
    static constexpr size_t kTotal = (kRing || kColRing) ? kBarOff + 8 * (size_t)D.nstage : kExchBytes;
};

// Groups 1.. of a plan whose input stage doubles as the second exchange buffer (FusedLayout::kStageExch): group g reads
// buffer A when g is odd and the stage when g is even, and writes the other one; the last group writes `dst`.  The
// stage is refilled (`recycle`) right after the barrier that follows its last reader.
template <class A, PlanDesc D, int g, class Env, class Dst, class Rec>
KF_HD void run_groups_stage_exch(Env& env, int t, bool active, const Dst& dst, typename A::C* exA, typename A::C* exS,
                                 const TwTab<A>& tw, const PlanConsts<A>& pc, int inverse, const Rec& recycle)
{
    constexpr bool kReadsStage = (g % 2 == 0);
    constexpr bool kLastStageReader = kReadsStage && (g + 2 > D.G - 1);
    run_group<A, D, g, NoSrc, Dst>(t, active, NoSrc{}, dst, kReadsStage ? exS : exA, kReadsStage ? exA : exS, tw, pc, inverse);
    if constexpr (g + 1 < D.G) {
        env.sync();
        if constexpr (kLastStageReader) recycle();
        run_groups_stage_exch<A, D, g + 1, Env, Dst, Rec>(env, t, active, dst, exA, exS, tw, pc, inverse, recycle);
    } else {
        env.sync();       // the stage (or buffer A, which the next tile's first group overwrites) has been consumed
        if constexpr (kLastStageReader) recycle();
    }
}

// PT is a tag type carrying the plan as `static constexpr PlanDesc D`.
// Env abstracts the execution environment (thread/block ids, CTA barrier, dynamic shared memory, bulk async
// copies + mbarriers) so that the identical body runs as a CUDA kernel (DeviceEnv, kf_kernels.cuh) and, for
// index-math tests without a GPU, under tests/emul's thread-per-CUDA-thread emulator.
//
// Input path when D.nstage > 0 (all modes with contiguous input rows): one elected thread keeps `nstage` tiles
// in flight with cp.async.bulk (global -> shared, completion on an mbarrier); the team reads its row from the
// landed stage with conflict-free LDS.  A stage is recycled right after the barrier that follows the only
// group that reads it, so the HBM read stream never waits for the butterflies.  With D.nstage == 0 (and always
// for the column mode, whose rows are strided) the first group loads straight from global memory.
template <class A, class PT, int MODE, class Env>
KF_HD void fused_body(const KParams<A>& P, Env& env)
{
    constexpr PlanDesc D = PT::D;
    typedef FusedLayout<A, PT, MODE> LY;
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    static_assert(D.valid(), "inconsistent plan descriptor");
    unsigned char* const smem_raw = env.smem();
    C* const smem = reinterpret_cast<C*>(smem_raw);
    constexpr int kPitch = D.pitch();
    constexpr bool kRing = LY::kRing;
    constexpr int kRowIn = LY::kRowIn;
    // two exchange buffers of tpc * pitch elements each
    C* const bufA = smem;
    C* const bufB = (D.nbuf == 2) ? smem + D.tpc * kPitch : smem;

    const int tid = env.tid();
    const int team = tid / D.team, t = tid % D.team;              // standard mapping: a team owns a transform
    // loop-invariant tables of this thread's work items (registers; PlanDesc::hoist)
    constexpr int kPairG = LY::kPairedLast ? D.G - 1 : -1;
    constexpr int kPairR = LY::kPairedLast ? D.R(D.G - 1) : (LY::kPairedFirst ? D.R(0) : 2);
    constexpr int kPairM = D.N / kPairR;
    constexpr bool kHoistSplit = (D.hoist & 1) && (LY::kPairedLast || LY::kPairedFirst);
    static_assert(!(D.hoist != 0 && LY::kShflPost), "hoisted tables assume the standard thread numbering");
    X hreg[D.hoist_total() > 0 && (D.hoist & 6) ? D.hoist_total() : 1];
    X hst[kHoistSplit ? ((kPairM / 2 + D.team - 1) / D.team) * kPairR : 1];
    {
        const TwTab<A> tw0{P.tw, P.gtw, P.g0tw, P.ctw, nullptr};
        if constexpr ((D.hoist & 6) != 0) hoist_fill<A, PT::D, kPairG>(t, tw0, hreg);
        if constexpr (kHoistSplit) hoist_fill_split<A, PT::D, kPairR, kPairM>(t, P.stw, hst);
    }
    const TwTab<A> tw{P.tw, P.gtw, P.g0tw, P.ctw, hreg};
    const long long ntiles = (P.howmany + D.tpc - 1) / D.tpc;
    int par = 0;   // parity of the exchange buffer sequence, carried across tiles (see run_groups)

    // ---- input ring ----
    auto stage_ptr = [&](int s) { return reinterpret_cast<C*>(smem_raw + LY::kRingOff + (size_t)s * LY::kStageBytes); };
    auto bar_ptr = [&](int s) { return smem_raw + LY::kBarOff + 8 * (size_t)s; };
    auto tile_mis = [&](long long tl) -> int {   // elements between the 16-byte grid and the start of tile tl
        if constexpr (LY::kSlackRing) return (int)((tl * PT::D.tpc * LY::kRowIn) % LY::kE16);
        else return 0;
    };
    auto issue = [&](int s, long long tl) {   // elected thread only
        if constexpr (LY::kColRing) {
            // tile tl = tpc adjacent columns of one plane, all N rows, as N / kBoxRows tensor-map boxes on one barrier
            const long long cb0 = tl * PT::D.tpc;
            const long long plane = P.ncols > 0 ? cb0 / P.ncols : 0, col0 = P.ncols > 0 ? P.in_col(cb0 % P.ncols) : cb0;
            env.mbar_expect(bar_ptr(s), (unsigned)((size_t)PT::D.tpc * PT::D.N * sizeof(C)));
            for (int r0 = 0; r0 < PT::D.N; r0 += LY::kBoxRows)
                env.tensor_load_box(bar_ptr(s), stage_ptr(s) + (size_t)r0 * PT::D.tpc, P.tmap, col0, r0, plane,
                                    P.in + plane * P.in_pdist + col0 + (long long)r0 * P.in_stride, P.in_stride, PT::D.tpc,
                                    LY::kBoxRows, (int)sizeof(C), r0 + LY::kBoxRows >= PT::D.N);
            return;
        }
        const long long row0 = tl * D.tpc;
        const long long rows = (P.howmany - row0) < D.tpc ? (P.howmany - row0) : D.tpc;
        if constexpr (LY::kSlackRing) {
            const int mis = tile_mis(tl);
            const long long elems = (rows * kRowIn + mis + LY::kE16 - 1) / LY::kE16 * LY::kE16;
            env.bulk_load(bar_ptr(s), stage_ptr(s), P.in + row0 * kRowIn - mis, (unsigned)(elems * sizeof(C)));
        } else {
            env.bulk_load(bar_ptr(s), stage_ptr(s), P.in + row0 * kRowIn, (unsigned)(rows * kRowIn * sizeof(C)));
        }
    };
    if constexpr (kRing || LY::kColRing) {
        if (tid == 0)
            for (int s = 0; s < D.nstage; ++s) env.mbar_init(bar_ptr(s));
        env.mbar_fence_init();
        env.sync();
        if (tid == 0)
            for (int s = 0; s < D.nstage; ++s) {
                const long long tl = env.bid() + (long long)s * env.nblocks();
                if (tl < ntiles) issue(s, tl);
            }
    }

    int it = 0;
    for (long long tile = env.bid(); tile < ntiles; tile += env.nblocks(), ++it) {
        const long long b = tile * D.tpc + team;
        const bool active = b < P.howmany;
        C* b0 = (par ? bufB : bufA) + team * kPitch;
        C* b1 = (par ? bufA : bufB) + team * kPitch;
        const int stg = (kRing || LY::kColRing) ? it % (D.nstage > 0 ? D.nstage : 1) : 0;
        const C* srow = nullptr;
        if constexpr (kRing) {
            env.mbar_wait(bar_ptr(stg), it / (D.nstage > 0 ? D.nstage : 1));
            srow = stage_ptr(stg) + tile_mis(tile) + team * kRowIn;
        } else if constexpr (LY::kColRing) {
            env.mbar_wait(bar_ptr(stg), it / (D.nstage > 0 ? D.nstage : 1));
            srow = stage_ptr(stg);
        }
        // after the CTA barrier that follows the last read of the stage: refill it with the tile nstage rounds ahead
        auto recycle = [&]() {
            if constexpr (LY::kRing || LY::kColRing) {
                if (tid == 0) {
                    const long long nxt = tile + (long long)PT::D.nstage * env.nblocks();
                    if (nxt < ntiles) issue(stg, nxt);
                }
            }
        };

        if constexpr (MODE == kR2C && LY::kStageExch) {
            // ---- kiss_fftr, three groups: stage -> A -> stage -> paired last group ----
            C* const exA = bufA + team * kPitch;
            C* const exS = stage_ptr(stg) + team * kPitch;
            DstGlobal<A> unused{nullptr};
            typedef SrcShared<A> S;
            run_group<A, PT::D, 0, S, DstGlobal<A>>(t, active, S{srow}, unused, nullptr, exA, tw, P.pc, P.inverse);
            env.sync();                                   // every team has consumed its landed row
            run_group<A, PT::D, 1, S, DstGlobal<A>>(t, active, S{srow}, unused, exA, exS, tw, P.pc, P.inverse);
            env.sync();
            run_r2c_last_paired<A, PT::D>(t, active, exS, tw, P.pc, P.stw, hst, P.out + b * P.out_dist, P.inverse, [&]() {
                env.sync();                               // the stage is consumed again: refill it
                recycle();
            });
        } else if constexpr (MODE == kR2C && LY::kPairedLast) {
            // ---- kiss_fftr: all groups but the last, then the last group + split pass on work-item pairs ----
            constexpr int gl = PT::D.G - 1;
            DstGlobal<A> unused{nullptr};
            auto run_all = [&](auto src) {
                typedef decltype(src) S;
                run_group<A, PT::D, 0, S, DstGlobal<A>>(t, active, src, unused, b0, b1, tw, P.pc, P.inverse);
                env.sync();
                recycle();
                run_groups<A, PT::D, 1, S, DstGlobal<A>, Env, gl>(env, t, active, src, unused, b1, b0, tw, P.pc, P.inverse);
            };
            if constexpr (kRing) run_all(SrcShared<A>{srow});
            else run_all(SrcGlobal<A, true>{P.in + b * P.in_dist, 1});
            run_r2c_last_paired<A, PT::D>(t, active, (gl & 1) ? b1 : b0, tw, P.pc, P.stw, hst, P.out + b * P.out_dist, P.inverse, []() {});
            par ^= (D.G - 1) & 1;
        } else if constexpr (MODE == kC2R && LY::kStageExch) {
            // ---- kiss_fftri, three groups: paired first group from the stage -> A -> stage -> last group ----
            C* const exA = bufA + team * kPitch;
            C* const exS = stage_ptr(stg) + team * kPitch;
            DstGlobal<A> dstg{P.out + b * P.out_dist};
            auto f = [&](int i) { return A::load(srow[i]); };
            run_c2r_first_paired<A, PT::D>(t, active, f, exA, tw, P.pc, P.stw, hst, P.inverse, []() {});
            env.sync();                                   // every team has consumed its landed row
            run_group<A, PT::D, 1, NoSrc, DstGlobal<A>>(t, active, NoSrc{}, dstg, exA, exS, tw, P.pc, P.inverse);
            env.sync();
            run_group<A, PT::D, 2, NoSrc, DstGlobal<A>>(t, active, NoSrc{}, dstg, exS, nullptr, tw, P.pc, P.inverse);
            env.sync();
            recycle();
        } else if constexpr (MODE == kC2R && LY::kPairedFirst) {
            // ---- kiss_fftri: split pre pass + first group on work-item pairs, then the remaining groups ----
            DstGlobal<A> dstg{P.out + b * P.out_dist};
            if constexpr (kRing) {
                auto f = [&](int i) { return A::load(srow[i]); };
                run_c2r_first_paired<A, PT::D>(t, active, f, b1, tw, P.pc, P.stw, hst, P.inverse, []() {});
            } else {
                const C* row = P.in + b * P.in_dist;
                auto f = [&](int i) { return A::load(ld_stream(row + i)); };
                run_c2r_first_paired<A, PT::D>(t, active, f, b1, tw, P.pc, P.stw, hst, P.inverse, []() {});
            }
            env.sync();
            recycle();
            run_groups<A, PT::D, 1, NoSrc, DstGlobal<A>>(env, t, active, NoSrc{}, dstg, b1, b0, tw, P.pc, P.inverse);
            par ^= (D.G - 1) & 1;
        } else if constexpr (MODE == kR2C && LY::kShflPost) {
            // ---- kiss_fftr with the split post pass (kiss_fftr.c:88-116) done in registers ---------------------------
            // The last group keeps its outputs T[k] in registers.  Output k needs T[k] and T[nc-k]; with M = N/R work
            // items, T[kp + j*M]'s partner is register R-1-j of work item M-kp.  The last group's threads are
            // renumbered so that the two owners of a pair sit in lanes l and 31-l of one warp, and the partner values
            // travel by warp shuffle -- no T[] buffer, no extra barrier.  Thread 0 and thread team/2 pair with
            // themselves.
            constexpr int gl = PT::D.G - 1, R = PT::D.R(gl), M = PT::D.items(gl), IT = PT::D.iters(gl), T = PT::D.team;
            constexpr int nc = PT::D.N;
            const int lane = tid & 31, wq = t >> 5;
            int tl = (lane < 16) ? 16 * wq + lane : T - (16 * wq + 31 - lane);
            if (tl == T) tl = T / 2;
            X treg[IT * R];
            DstRegs<A, R> dstr{treg};
            auto run_all = [&](auto src) {
                typedef decltype(src) S;
                if constexpr (PT::D.G == 1) {
                    run_group<A, PT::D, 0, S, DstRegs<A, R>>(tl, active, src, dstr, b0, b1, tw, P.pc, P.inverse);
                    if constexpr (LY::kRing) env.sync();
                    recycle();
                } else {
                    run_group<A, PT::D, 0, S, DstRegs<A, R>>(t, active, src, dstr, b0, b1, tw, P.pc, P.inverse);
                    env.sync();
                    recycle();
                    // middle groups, then the last one under the pairing-friendly thread numbering
                    run_groups<A, PT::D, 1, S, DstRegs<A, R>, Env, gl>(env, t, active, src, dstr, b1, b0, tw, P.pc, P.inverse);
                    run_group<A, PT::D, gl, S, DstRegs<A, R>>(tl, active, src, dstr, (gl & 1) ? b1 : b0, nullptr, tw, P.pc, P.inverse);
                }
            };
            if constexpr (kRing) run_all(SrcShared<A>{srow});
            else run_all(SrcGlobal<A, true>{P.in + b * P.in_dist, 1});
            if (active) {
                const bool is0 = (tl == 0), self = is0 || (tl == T / 2);
                const int srcl = self ? lane : 31 - lane;
                C* out = P.out + b * P.out_dist;
                static_for<IT>([&](auto ITER) {
                    static_for<R>([&](auto E) {
                        constexpr int it = decltype(ITER)::value, e = decltype(E)::value;
                        constexpr int j = PT::D.kout(gl, e) / M;
                        constexpr int pe = PT::D.reg_of_j(gl, R - 1 - j), pit = IT - 1 - it;           // general partner
                        constexpr int p0e = (it == 0) ? PT::D.reg_of_j(gl, (R - j) % R) : pe;        // thread 0's partner
                        constexpr int p0it = (IT - it) % IT;
                        X pg{env.shfl(treg[pit * R + pe].r, srcl), env.shfl(treg[pit * R + pe].i, srcl)};
                        const X partner = is0 ? treg[p0it * R + p0e] : pg;
                        const X own = treg[it * R + e];
                        const int k = tl + it * T + j * M;
                        if (it == 0 && j == 0 && is0) {
                            X ok, onk;
                            fftr_post_pair<A>(0, nc, own, own, own, ok, onk);      // DC and Nyquist bins
                            out[0] = A::store(ok);
                            out[nc] = A::store(onk);
                        } else {
                            const bool lo = 2 * k < nc;             // k == nc/2: the reference's second assignment wins
                            const int ks = lo ? k : nc - k;
                            const X st = A::load(TwTab<A>::ro_load_c(P.stw + (ks - 1)));
                            X ok, onk;
                            fftr_post_pair<A>(ks, nc, lo ? own : partner, lo ? partner : own, st, ok, onk);
                            out[k] = A::store(lo ? ok : onk);
                        }
                    });
                });
            }
            if constexpr (D.G >= 2) par ^= (D.G - 1) & 1;
        } else if constexpr (MODE == kC2C && LY::kStageExch) {
            // ---- C2C, three or more groups: stage -> A -> stage -> ... -> global ----
            C* const exA = bufA + team * kPitch;
            C* const exS = stage_ptr(stg) + team * kPitch;
            DstGlobal<A> dstg{P.out + b * P.out_dist};
            run_group<A, PT::D, 0, SrcShared<A>, DstGlobal<A>>(t, active, SrcShared<A>{srow}, dstg, nullptr, exA, tw, P.pc, P.inverse);
            env.sync();                                   // every team has consumed its landed row
            run_groups_stage_exch<A, PT::D, 1>(env, t, active, dstg, exA, exS, tw, P.pc, P.inverse, recycle);
        } else if constexpr (MODE == kC2C || MODE == kR2C || MODE == kC2R) {
            // kR2C: the real row is read as ncfft packed complex (kiss_fftr.c:77); the last group leaves T[] in
            // natural order in the next exchange buffer for the split pass
            C* tb = (((D.G - 1) & 1) ? b0 : b1);
            DstGlobal<A> dstg{P.out + b * P.out_dist};
            DstShared<A> dsts{tb};
            auto run_from = [&](auto src, auto dst) {
                typedef decltype(src) S;
                typedef decltype(dst) Dd;
                run_group<A, PT::D, 0, S, Dd>(t, active, src, dst, b0, b1, tw, P.pc, P.inverse);
                if constexpr (PT::D.G >= 2 || LY::kRing) env.sync();
                recycle();
                if constexpr (PT::D.G >= 2) run_groups<A, PT::D, 1, S, Dd>(env, t, active, src, dst, b1, b0, tw, P.pc, P.inverse);
            };
            auto with_dst = [&](auto src) {
                if constexpr (MODE == kR2C) run_from(src, dsts);
                else run_from(src, dstg);
            };
            if constexpr (MODE == kC2R) {
                // kiss_fftri: the split pre pass is evaluated inside the first group's loads (SrcC2RFused)
                if constexpr (kRing) with_dst(SrcC2RFused<A, true>{srow, P.stw, PT::D.N});
                else with_dst(SrcC2RFused<A, false>{P.in + b * P.in_dist, P.stw, PT::D.N});
            } else if constexpr (kRing) with_dst(SrcShared<A>{srow});
            else with_dst(SrcGlobal<A, true>{P.in + b * P.in_dist, 1});   // contiguous rows (other strides: generic kernel)
            if constexpr (MODE != kR2C) {
                if constexpr (D.G >= 2) par ^= (D.G - 1) & 1;
            } else {
                env.sync();
                if (active) {
                    // split post pass (kiss_fftr.c:88-116), one bin pair (k, nc-k) per step; fully unrolled so the
                    // shared-memory reads, twiddle fetches and stores of different pairs overlap
                    constexpr int nc = PT::D.N, kPairs = nc / 2 + 1, kIt = (kPairs + PT::D.team - 1) / PT::D.team;
                    C* out = P.out + b * P.out_dist;
                    static_for<kIt>([&](auto I) {
                        const int k = t + decltype(I)::value * PT::D.team;
                        if ((decltype(I)::value + 1) * PT::D.team <= kPairs || k < kPairs) {
                            X Tk = A::load(tb[k]);
                            X Tnk = (k == 0) ? Tk : A::load(tb[nc - k]);
                            X st = (k == 0) ? Tk : A::load(TwTab<A>::ro_load_c(P.stw + (k - 1)));
                            X ok, onk;
                            fftr_post_pair<A>(k, nc, Tk, Tnk, st, ok, onk);
                            out[k] = A::store(ok);
                            out[nc - k] = A::store(onk);
                        }
                    });
                }
                par ^= D.G & 1;   // G exchanges were used (G-1 between groups + the T[] buffer)
            }
        } else if constexpr (MODE == kC2CCol) {
            // group 0 uses the transposed mapping: consecutive lanes take consecutive transforms (columns)
            static_assert(D.G >= 2, "column mode needs a shared-memory exchange");
            const int cteam = tid % D.tpc, ct = tid / D.tpc;
            const long long cb = tile * D.tpc + cteam;
            SrcGlobal<A, false> src{P.in + P.in_off(cb), P.in_stride};
            DstGlobal<A> dst{P.out_ptr(b)};
            C* cb1 = (par ? bufA : bufB) + cteam * kPitch;
            if constexpr (LY::kColRing) {
                SrcStageCol<A> ssrc{srow + cteam, PT::D.tpc};
                run_group<A, D, 0, SrcStageCol<A>, DstGlobal<A>>(ct, cb < P.howmany, ssrc, dst, nullptr, cb1, tw, P.pc, P.inverse);
            } else {
                run_group<A, D, 0, SrcGlobal<A, false>, DstGlobal<A>>(ct, cb < P.howmany, src, dst, nullptr, cb1, tw, P.pc, P.inverse);
            }
            env.sync();
            recycle();                                    // the landed tile has been consumed: fetch the one nstage rounds ahead
            run_groups<A, D, 1>(env, t, active, src, dst, b1, b0, tw, P.pc, P.inverse);
            par ^= (D.G - 1) & 1;
        }
        else if constexpr (MODE == kC2CColTw || MODE == kC2CColCol) {
            // ---- four-step transform of a long row (float / double), see kf_api.c:kf_exec_fourstep ----
            static_assert(D.G >= 2 && (MODE == kC2CColCol || !A::kFixed), "column modes need a shared-memory exchange; the twiddled pass is float / double only");
            const int cteam = tid % D.tpc, ct = tid / D.tpc;
            const long long cb = tile * D.tpc + cteam;
            const bool con = cb < P.howmany;
            constexpr int gl = PT::D.G - 1;
            SrcGlobal<A, false> src{P.in + P.in_off(cb), P.in_stride};
            DstGlobal<A> unused{nullptr};
            C* cb1 = (par ? bufA : bufB) + cteam * kPitch;
            C* cb0 = (par ? bufB : bufA) + cteam * kPitch;
            if constexpr (LY::kColRing) {
                SrcStageCol<A> ssrc{srow + cteam, PT::D.tpc};
                run_group<A, D, 0, SrcStageCol<A>, DstGlobal<A>>(ct, con, ssrc, unused, nullptr, cb1, tw, P.pc, P.inverse);
            } else {
                run_group<A, D, 0, SrcGlobal<A, false>, DstGlobal<A>>(ct, con, src, unused, nullptr, cb1, tw, P.pc, P.inverse);
            }
            env.sync();
            recycle();
            run_groups<A, D, 1, SrcGlobal<A, false>, DstGlobal<A>, Env, gl>(env, t, active, src, unused, b1, b0, tw, P.pc, P.inverse);
            if constexpr (MODE == kC2CColTw) {
                // standard mapping: the team writes its transform as a contiguous row, times W_N^(column * k)
                const long long col = P.ncols > 0 ? b % P.ncols : b;
                DstGlobalTw<A> dst{P.out_ptr(b), P.stw, col};
                run_group<A, D, gl, SrcGlobal<A, false>, DstGlobalTw<A>>(t, active, src, dst, (gl & 1) ? b1 : b0, nullptr, tw, P.pc, P.inverse);
            } else {
                // transposed mapping again: adjacent lanes store adjacent columns, element k lands k * in_stride below
                DstGlobalStrided<A> dst{P.out + P.out_off(cb), P.in_stride};
                run_group<A, D, gl, SrcGlobal<A, false>, DstGlobalStrided<A>>(ct, con, src, dst, (gl & 1) ? cb1 : cb0, nullptr, tw, P.pc, P.inverse);
            }
            par ^= (D.G - 1) & 1;
        }
        // single exchange buffer: the next tile's first group overwrites what the last group just read
        if constexpr (D.nbuf == 1 && !LY::kStageExch) env.sync();
    }
}

// =========================================================================================================
// Fused fast convolution (overlap-scrap), the step either side of the transform in the reference's own flagship
// tool (tools/kiss_fastfir.c:167-181 fastconv1buf, :191-206 kff_nocopy): for every block b of nfft input samples
// starting at b*ngood:   out[b*ngood .. +ngood) = IFFT( FFT(in block) .* H )[0 .. ngood).
// Three launches and two HBM round trips of the spectrum in the reference's structure; here ONE kernel: the
// forward plan's last group keeps the spectrum in registers, multiplies by H, and hands the registers to the first
// group of the inverse plan (same PlanDesc, conjugate twiddles) -- possible because the plan is palindromic
// (R(last group) == R(first group)), so thread kp's forward outputs {kp + j*M} are exactly its inverse inputs.
// HBM traffic per block: nfft reads + ngood writes (+ the L2-resident H and twiddle tables).
// =========================================================================================================
template <class A>
struct FCSide {                       // one direction's tables (see KParams)
    const typename A::C* tw;
    const typename A::C* gtw;
    cx<typename A::R> g0tw[kMaxG0Slots];
    cx<typename A::R> ctw[kMaxCtw];
    PlanConsts<A> pc;
};

template <class A>
struct FCParams {
    const typename A::C* in;
    typename A::C* out;
    long long nblocks;
    long long ngood;                  // block advance on both sides, and the number of samples stored per block
    const typename A::C* h;           // nfft-point frequency response, already scaled by 1/nfft (kiss_fastfir.c:149-162)
    FCSide<A> fwd, inv;
};

template <class A, int R_, int GL_, PlanDesc D>
struct SrcRegs {                      // inverse group 0 input element off + e*Flo == forward output kp + e*M
    const cx<typename A::R>* regs;
    template <int IT_, int E_>
    KF_HD cx<typename A::R> get(int) const { return regs[IT_ * R_ + D.reg_of_j(GL_, E_)]; }
};

template <class A>
struct DstGlobalClip {
    typename A::C* base;
    int ngood;
    template <int IT_, int E_>
    KF_HD void put(int k, const cx<typename A::R>& v) const
    {
        if (k < ngood) base[k] = A::store(v);
    }
};

template <class A, class PT, class Env>
KF_HD void fastconv_body(const FCParams<A>& P, Env& env)
{
    constexpr PlanDesc D = PT::D;
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    static_assert(!A::kFixed, "fast convolution is float/double only (reference: kiss_fastfir.c:152)");
    static_assert(D.R(0) == D.R(D.G - 1) && D.iters(0) == D.iters(D.G - 1), "fast convolution needs a palindromic plan");
    static_assert(D.hoist == 0, "the fast-convolution kernel does not hoist tables");
    constexpr int GL = D.G - 1, R = D.R(GL), IT = D.iters(GL), M = D.items(GL);
    constexpr int kPitch = D.pitch();
    C* const smem = reinterpret_cast<C*>(env.smem());
    C* const bufA = smem;
    C* const bufB = smem + D.tpc * kPitch;
    const int tid = env.tid();
    const int team = tid / D.team, t = tid % D.team;
    const TwTab<A> twf{P.fwd.tw, P.fwd.gtw, P.fwd.g0tw, P.fwd.ctw};
    const TwTab<A> twi{P.inv.tw, P.inv.gtw, P.inv.g0tw, P.inv.ctw};
    const long long ntiles = (P.nblocks + D.tpc - 1) / D.tpc;
    // exchanges per tile: 2*(G-1), always even -> the ping-pong parity is the same for every tile
    for (long long tile = env.bid(); tile < ntiles; tile += env.nblocks()) {
        const long long b = tile * D.tpc + team;
        const bool active = b < P.nblocks;
        C* b0 = bufA + team * kPitch;
        C* b1 = bufB + team * kPitch;
        X spec[IT * R];
        // ---- forward transform, spectrum left in registers ----
        SrcGlobal<A, true> src{P.in + b * P.ngood, 1};
        DstRegs<A, R> dstr{spec};
        if constexpr (D.G == 1) {
            run_group<A, D, 0, SrcGlobal<A, true>, DstRegs<A, R>>(t, active, src, dstr, b0, b1, twf, P.fwd.pc, 0);
        } else {
            run_groups<A, D, 0, SrcGlobal<A, true>, DstRegs<A, R>, Env, D.G>(env, t, active, src, dstr, b0, b1, twf, P.fwd.pc, 0);
        }
        // ---- pointwise multiply by the frequency response (fastconv1buf, kiss_fastfir.c:171-176) ----
        if (active) {
            static_for<IT>([&](auto ITER) {
                static_for<R>([&](auto E) {
                    constexpr int it = decltype(ITER)::value, e = decltype(E)::value;
                    const int k = t + it * PT::D.team + PT::D.kout(GL, e);
                    if ((it + 1) * PT::D.team <= M || t + it * PT::D.team < M) {
                        const X hk = A::load(TwTab<A>::ro_load_c(P.h + k));
                        spec[it * R + e] = A::cmul(spec[it * R + e], hk);
                    }
                });
            });
        }
        // ---- inverse transform from the registers; only the ngood valid samples are stored ----
        SrcRegs<A, R, GL, D> srcr{spec};
        DstGlobalClip<A> dstc{P.out + b * P.ngood, (int)P.ngood};
        // the forward pass used G-1 exchanges: continue the ping-pong where it stopped
        C* r0 = ((D.G - 1) & 1) ? b1 : b0;
        C* r1 = ((D.G - 1) & 1) ? b0 : b1;
        if constexpr (D.G == 1) {
            run_group<A, D, 0, SrcRegs<A, R, GL, D>, DstGlobalClip<A>>(t, active, srcr, dstc, r0, r1, twi, P.inv.pc, 1);
        } else {
            run_groups<A, D, 0, SrcRegs<A, R, GL, D>, DstGlobalClip<A>, Env, D.G>(env, t, active, srcr, dstc, r0, r1, twi, P.inv.pc, 1);
        }
    }
}

// =========================================================================================================
// Run-time plan kernel: any radix schedule, one stage per shared-memory pass.
// =========================================================================================================
struct GenericPlan {
    int N, L;
    int p[32], m[32];   // kf_factor order (kiss_fft.c:306-328), MAXFACTORS == 32 (_kiss_fft_guts.h:21)
};

template <class A>
struct GParams {
    KParams<A> k;
    GenericPlan plan;
    int mode;   // Mode
    int tpc;    // transforms per CTA (kC2CCol: adjacent columns)
};

// One output element of the generic radix (kiss_fft.c:192-233) -- every thread owns one output so that even a
// prime nfft (single stage p == nfft) is spread over the whole CTA.
template <class A>
KF_HD cx<typename A::R> generic_output(const typename A::C* rd, int u, int off, int q1, int p, int m,
                                                            int F, int N, const TwTab<A>& tw)
{
    typedef cx<typename A::R> X;
    const int k = u + q1 * m;
    const int step = (int)(((long long)F * k) % N);
    int twidx = 0;
    X s0 = A::load(rd[(u * p) * F + off]);
    X acc{A::divk_rt(s0.r, p), A::divk_rt(s0.i, p)};
    for (int q = 1; q < p; ++q) {
        twidx += step;
        if (twidx >= N) twidx -= N;
        X sq = A::load(rd[(u * p + q) * F + off]);
        sq = X{A::divk_rt(sq.r, p), A::divk_rt(sq.i, p)};
        acc = cadd<A>(acc, A::cmul_bf(sq, tw.get(twidx)));
    }
    return cwrap<A>(acc);
}

template <class A, class Env>
KF_HD void generic_body(const GParams<A>& G, Env& env)
{
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    const KParams<A>& P = G.k;
    const int N = G.plan.N, L = G.plan.L, tpc = G.tpc, mode = G.mode;
    C* const buf0 = reinterpret_cast<C*>(env.smem());
    C* const buf1 = buf0 + (size_t)tpc * N;
    const TwTab<A> tw{P.tw, nullptr, nullptr, nullptr};
    const int nthr = env.nthreads(), tid = env.tid();
    const long long ntiles = (P.howmany + tpc - 1) / tpc;

    for (long long tile = env.bid(); tile < ntiles; tile += env.nblocks()) {
        const long long bbase = tile * tpc;
        const int nb = (int)((P.howmany - bbase) < tpc ? (P.howmany - bbase) : tpc);
        // ---- load (level L, natural order) ----
        if (mode == kC2R) {
            for (int i = tid; i < nb * (N / 2 + 1); i += nthr) {
                const int bl = i / (N / 2 + 1), k = i % (N / 2 + 1);
                const C* in = P.in + (bbase + bl) * P.in_dist;
                X Fk = A::load(in[k]), Fnk = A::load(in[N - k]);
                X st = (k == 0) ? Fk : A::load(P.stw[k - 1]);
                X Tk, Tnk;
                fftri_pre_pair<A>(k, Fk, Fnk, st, Tk, Tnk);
                // k == N-k (N even, k == N/2): the reference's second assignment wins (kiss_fftr.c:147-153)
                if (k != N - k) buf0[bl * N + k] = A::store(Tk);
                if (k != 0) buf0[bl * N + (N - k)] = A::store(Tnk);
            }
        } else if (mode == kC2CCol) {
            for (int i = tid; i < tpc * N; i += nthr) {
                const int bl = i % tpc, n = i / tpc;
                if (bl < nb) buf0[bl * N + n] = P.in[P.in_off(bbase + bl) + (long long)n * P.in_stride];
            }
        } else {
            for (int i = tid; i < nb * N; i += nthr) {
                const int bl = i / N, n = i % N;
                buf0[i] = P.in[(bbase + bl) * P.in_dist + (long long)n * P.in_stride];
            }
        }
        env.sync();
        // ---- radix stages, innermost first ----
        C* rd = buf0;
        C* wr = buf1;
        int F = N;
        for (int s = L - 1; s >= 0; --s) {
            const int p = G.plan.p[s], m = G.plan.m[s];
            F /= p;
            if (p == 2 || p == 3 || p == 4 || p == 5) {
                const int nbf = N / p;
                for (int i = tid; i < nb * nbf; i += nthr) {
                    const int bl = i / nbf, j = i % nbf;
                    const int k = j / F, off = j % F;
                    const C* r = rd + bl * N;
                    C* w = wr + bl * N;
                    X v[5];
                    for (int q = 0; q < p; ++q) v[q] = A::load(r[(k * p + q) * F + off]);
                    if (p == 2) bfly2<A, false>(v, tw.get(F * k));
                    else if (p == 4) bfly4<A, false>(v, tw.get(F * k), tw.get(2 * F * k), tw.get(3 * F * k), A::sign_of(P.inverse));
                    else if (p == 3) bfly3<A, false>(v, tw.get(F * k), tw.get(2 * F * k), P.pc.epi3.i);
                    else bfly5<A, false>(v, tw.get(F * k), tw.get(2 * F * k), tw.get(3 * F * k), tw.get(4 * F * k), P.pc.ya, P.pc.yb);
                    for (int q = 0; q < p; ++q) w[(k + q * m) * F + off] = A::store(v[q]);
                }
            } else {
                for (int i = tid; i < nb * N; i += nthr) {
                    const int bl = i / N, o = i % N;          // o = output address (k*F + off)
                    const int k = o / F, off = o % F;
                    const int q1 = k / m, u = k % m;
                    wr[bl * N + o] = A::store(generic_output<A>(rd + bl * N, u, off, q1, p, m, F, N, tw));
                }
            }
            env.sync();
            C* tsw = rd; rd = wr; wr = tsw;
        }
        // ---- store (level 0, natural order, in rd) ----
        if (mode == kR2C) {
            for (int i = tid; i < nb * (N / 2 + 1); i += nthr) {
                const int bl = i / (N / 2 + 1), k = i % (N / 2 + 1);
                const C* T = rd + bl * N;
                C* out = P.out + (bbase + bl) * P.out_dist;
                X Tk = A::load(T[k]);
                X Tnk = (k == 0) ? Tk : A::load(T[N - k]);
                X st = (k == 0) ? Tk : A::load(P.stw[k - 1]);
                X ok, onk;
                fftr_post_pair<A>(k, N, Tk, Tnk, st, ok, onk);
                if (k != N - k) out[k] = A::store(ok);
                out[N - k] = A::store(onk);
            }
        } else {
            for (int i = tid; i < nb * N; i += nthr) {
                const int bl = i / N, k = i % N;
                P.out_ptr(bbase + bl)[k] = rd[i];
            }
        }
        env.sync();
    }
}

// =========================================================================================================
// Multi-pass path for lengths whose exchange buffers do not fit in shared memory: one radix stage of the same
// Stockham recurrence per launch, levels kept in global memory (autosort layout, dense rows of N elements in the
// work buffers; the first stage reads the caller's rows with their stride/distance, the last stage writes the
// caller's output rows).  Slow (2*L passes over HBM) but it makes every nfft the reference accepts work.
// =========================================================================================================
template <class A>
struct StageParams {
    const typename A::C* in;
    typename A::C* out;
    long long batch, in_dist, out_dist, in_stride;   // only used by the first (in_*) / last (out_dist) stage
    int N, p, m, F;                                  // stage: radix p, span m, twiddle stride F = prod of outer radices
    int first, last;
    const typename A::C* tw;
    PlanConsts<A> pc;
    int inverse;
};

template <class A, class Env>
KF_HD void stage_body(const StageParams<A>& S, Env& env)
{
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    const TwTab<A> tw{S.tw, nullptr, nullptr, nullptr};
    const int N = S.N, p = S.p, m = S.m, F = S.F;
    const bool small = (p == 2 || p == 3 || p == 4 || p == 5);
    const long long per = small ? N / p : N;
    const long long total = S.batch * per;
    const long long step = env.nblocks() * env.nthreads();
    for (long long i = env.bid() * env.nthreads() + env.tid(); i < total; i += step) {
        const long long b = i / per;
        const int j = (int)(i % per);
        const C* in = S.first ? S.in + b * S.in_dist : S.in + b * N;
        const long long istr = S.first ? S.in_stride : 1;
        C* out = S.last ? S.out + b * S.out_dist : S.out + b * N;
        if (small) {
            const int k = j / F, off = j % F;
            X v[5];
            for (int q = 0; q < p; ++q) v[q] = A::load(in[(long long)((k * p + q) * F + off) * istr]);
            if (p == 2) bfly2<A, false>(v, tw.get(F * k));
            else if (p == 4) bfly4<A, false>(v, tw.get(F * k), tw.get(2 * F * k), tw.get(3 * F * k), A::sign_of(S.inverse));
            else if (p == 3) bfly3<A, false>(v, tw.get(F * k), tw.get(2 * F * k), S.pc.epi3.i);
            else bfly5<A, false>(v, tw.get(F * k), tw.get(2 * F * k), tw.get(3 * F * k), tw.get(4 * F * k), S.pc.ya, S.pc.yb);
            for (int q = 0; q < p; ++q) out[(k + q * m) * F + off] = A::store(v[q]);
        } else {
            // one output of kf_bfly_generic (kiss_fft.c:192-233) per thread
            const int k = j / F, off = j % F;
            const int q1 = k / m, u = k % m;
            const int stepw = (int)(((long long)F * (u + q1 * m)) % N);
            int twidx = 0;
            X s0 = A::load(in[(long long)((u * p) * F + off) * istr]);
            X acc{A::divk_rt(s0.r, p), A::divk_rt(s0.i, p)};
            for (int q = 1; q < p; ++q) {
                twidx += stepw;
                if (twidx >= N) twidx -= N;
                X sq = A::load(in[(long long)((u * p + q) * F + off) * istr]);
                sq = X{A::divk_rt(sq.r, p), A::divk_rt(sq.i, p)};
                acc = cadd<A>(acc, A::cmul_bf(sq, tw.get(twidx)));
            }
            out[j] = A::store(cwrap<A>(acc));
        }
    }
}

// split passes of the real transforms as stand-alone passes (multi-pass path only)
template <class A>
struct RealPassParams {
    const typename A::C* in;
    typename A::C* out;
    long long batch, in_dist, out_dist;
    int nc;          // packed complex length
    int post;        // 1: kiss_fftr post pass T[nc] -> F[nc+1]; 0: kiss_fftri pre pass F[nc+1] -> T[nc]
    const typename A::C* stw;
};

template <class A, class Env>
KF_HD void realpass_body(const RealPassParams<A>& S, Env& env)
{
    typedef typename A::C C;
    typedef cx<typename A::R> X;
    const int nc = S.nc, pairs = nc / 2 + 1;
    const long long total = S.batch * pairs;
    const long long step = env.nblocks() * env.nthreads();
    for (long long i = env.bid() * env.nthreads() + env.tid(); i < total; i += step) {
        const long long b = i / pairs;
        const int k = (int)(i % pairs);
        const C* in = S.in + b * S.in_dist;
        C* out = S.out + b * S.out_dist;
        if (S.post) {
            X Tk = A::load(in[k]);
            X Tnk = (k == 0) ? Tk : A::load(in[nc - k]);
            X st = (k == 0) ? Tk : A::load(S.stw[k - 1]);
            X ok, onk;
            fftr_post_pair<A>(k, nc, Tk, Tnk, st, ok, onk);
            if (k != nc - k) out[k] = A::store(ok);
            out[nc - k] = A::store(onk);
        } else {
            X Fk = A::load(in[k]), Fnk = A::load(in[nc - k]);
            X st = (k == 0) ? Fk : A::load(S.stw[k - 1]);
            X Tk, Tnk;
            fftri_pre_pair<A>(k, Fk, Fnk, st, Tk, Tnk);
            if (k != nc - k) out[k] = A::store(Tk);
            if (k != 0) out[nc - k] = A::store(Tnk);
        }
    }
}

}   // namespace kf
