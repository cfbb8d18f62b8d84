// kf_plan.h -- compile-time description of one batched Stockham plan.
//
// The radix schedule is the reference's (kf_factor, kiss_fft.c:306-328): stage 0 is the outermost /
// last-executed radix, stage L-1 the innermost / first-executed one.  With
//     F_s = p_0 * ... * p_{s-1}     (twiddle stride "fstride" of stage s, kiss_fft.c:235-300)
//     m_s = N / (p_0 * ... * p_s)   (butterfly span of stage s)
// the recursion of kf_work is equivalent to the breadth-first recurrence
//     Y_s[off][k + r*m_s] = butterfly_{p_s}( Y_{s+1}[off + q*F_s][k] , twiddle tw[q*k*F_s] ),  q,r < p_s
// with Y_L[off][0] = x[off] and X[k] = Y_0[0][k].  Every level is stored with the autosort address
//     addr_s(off, k) = k*F_s + off
// which is the natural input order at level L and the natural output order at level 0, so no digit
// reversal pass exists anywhere.
//
// Consecutive stages are fused into "groups" (super-stages) that one thread executes entirely in registers:
// group g covers stages s_hi(g) >= s >= s_lo(g), holds R(g) = prod p_s elements per work item and there are
// N/R(g) work items per transform.  Groups exchange data through shared memory (level s_lo(g) array).
#pragma once
#include "kf_math.h"

#if defined(__CUDACC__)
#define KF_CE __host__ __device__ constexpr
#else
#define KF_CE constexpr
#endif

namespace kf {

constexpr int kMaxStages = 16;
constexpr int kMaxGroups = 8;
constexpr int kMaxG0Slots = 64;   // stage twiddles of group 0 carried in the kernel parameters
constexpr int kMaxCtw = 96;       // split-twiddle constants carried in the kernel parameters

struct PlanDesc {
    int N;                  // transform length
    int L;                  // number of radix stages
    int p[kMaxStages];      // p[0] outermost ... p[L-1] innermost (reference order)
    int G;                  // number of register groups; group 0 runs first and covers the innermost stages
    int glen[kMaxGroups];   // stages per group, sum == L
    int team;               // threads cooperating on one transform
    int tpc;                // transforms per CTA
    int logpad;             // shared-memory skew: phys(a) = a + (a >> logpad); >= 31 disables it
    int minblocks;          // __launch_bounds__ min CTAs per SM
    int nstage;             // depth of the bulk-async (TMA) input ring; 0 = first group loads directly from global
    int twmode;             // 0: every stage twiddle is fetched from the tables (exact copies of the reference's entries; always
                            //    the case in fixed point); 1 (float/double only): a butterfly whose already-transformed upper
                            //    digits are non-zero derives its twiddles as tw[q*F*k'] * tw[q*F*kabove] -- the first factor is the
                            //    table entry of the upper==0 butterfly (shared by the whole stage), the second a plan constant
    int shfl_post;          // kiss_fftr: split post pass in registers via warp shuffles (needs team % 32 == 0)
    int paired;             // real transforms on work-item pairs.  kiss_fftr plan: the last group runs items k' and m-k' in the
                            // same thread, so every bin pair (k, nc-k) of the split post pass is complete in its registers;
                            // kiss_fftri plan: the first group runs items u and W-u, reading every spectrum pair once
    int hoist;              // loop-invariant tables kept in registers across the tiles of a persistent CTA (a thread always runs
                            // the same work items): bit 0 = split twiddles of the paired real pass, bit 1 = stage twiddles of
                            // the unpaired groups >= 1, bit 2 = stage twiddles of the paired group
    int nbuf;               // exchange buffers: 2 = ping-pong (default); 1 = single buffer: two-group plans in the C2C /
                            // column modes (+ one barrier per tile; halves shared memory => wider column tiles), or
                            // three-group paired real plans, where the input stage doubles as the second buffer

    KF_CE int F(int s) const
    {
        int f = 1;
        for (int j = 0; j < s; ++j) f *= p[j];
        return f;
    }
    KF_CE int m(int s) const { return N / (F(s) * p[s]); }
    KF_CE int s_hi(int g) const
    {
        int s = L - 1;
        for (int j = 0; j < g; ++j) s -= glen[j];
        return s;
    }
    KF_CE int s_lo(int g) const { return s_hi(g) - glen[g] + 1; }
    KF_CE int R(int g) const
    {
        int r = 1;
        for (int s = s_lo(g); s <= s_hi(g); ++s) r *= p[s];
        return r;
    }
    KF_CE int Flo(int g) const { return F(s_lo(g)); }
    KF_CE int mhi(int g) const { return m(s_hi(g)); }
    KF_CE int items(int g) const { return N / R(g); }
    KF_CE int iters(int g) const { return (items(g) + team - 1) / team; }
    // weight of stage s's digit inside the register index of group g: W_{s_lo} = 1, W_{s+1} = W_s * p_s
    KF_CE int W(int g, int s) const
    {
        int w = 1;
        for (int j = s_lo(g); j < s; ++j) w *= p[j];
        return w;
    }
    KF_CE int digit(int g, int s, int e) const { return (e / W(g, s)) % p[s]; }
    // k offset contributed by the (already transformed) digits of stages above s inside group g
    KF_CE int kabove(int g, int s, int e) const
    {
        int k = 0;
        for (int j = s + 1; j <= s_hi(g); ++j) k += digit(g, j, e) * m(j);
        return k;
    }
    // output position offset (in units of k) of register e after the whole group ran
    KF_CE int kout(int g, int e) const
    {
        int k = 0;
        for (int j = s_lo(g); j <= s_hi(g); ++j) k += digit(g, j, e) * m(j);
        return k;
    }
    // ---- per-group twiddle tables ------------------------------------------------------------------------
    // Every butterfly (g, s, e) of a work item needs the p_s - 1 stage twiddles tw[q*F_s*(k' + kabove)].  They
    // depend on the work item only through k', so each (butterfly, q) pair gets a "slot" and the plan carries a
    // table gtw[g][slot][w] laid out with the work item w innermost: a warp's fetch of one slot is one contiguous,
    // immediate-addressed vector load instead of a gather over the N-entry twiddle array.  Group 0 has k' == 0:
    // its slots are plan constants and travel in the kernel parameters (constant bank).
    // butterflies of stage s that agree in the digits of the stages above s (already transformed inside the group)
    // have the same kabove and therefore the same twiddles: one slot set per combination of upper digits
    KF_CE int nupper(int g, int s) const { return R(g) / (W(g, s) * p[s]); }
    KF_CE int nslots_stage(int g, int s) const { return nupper(g, s) * (p[s] - 1); }
    KF_CE int slot(int g, int s, int e) const   // first slot of the butterfly whose base register is e
    {
        int n = 0;
        for (int j = s_hi(g); j > s; --j) n += nslots_stage(g, j);
        return n + (e / (W(g, s) * p[s])) * (p[s] - 1);
    }
    KF_CE int nslots(int g) const
    {
        int n = 0;
        for (int s = s_lo(g); s <= s_hi(g); ++s) n += nslots_stage(g, s);
        return n;
    }
    KF_CE int gtw_offset(int g) const   // entries before group g's table (groups 1..G-1 only)
    {
        int n = 0;
        for (int j = 1; j < g; ++j) n += nslots(j) * items(j);
        return n;
    }
    KF_CE int gtw_total() const { return gtw_offset(G); }
    // constants tw[q*F_s*kabove] of the split-twiddle mode: one per (group >= 1, stage, non-zero upper combination, q)
    KF_CE int nctw_stage(int g, int s) const { return (nupper(g, s) - 1) * (p[s] - 1); }
    KF_CE int cslot(int g, int s, int up, int q) const   // up >= 1, q >= 1
    {
        int n = 0;
        for (int gg = 1; gg < g; ++gg)
            for (int j = s_lo(gg); j <= s_hi(gg); ++j) n += nctw_stage(gg, j);
        for (int j = s_hi(g); j > s; --j) n += nctw_stage(g, j);
        return n + (up - 1) * (p[s] - 1) + (q - 1);
    }
    KF_CE int nctw() const
    {
        int n = 0;
        for (int gg = 1; gg < G; ++gg)
            for (int j = s_lo(gg); j <= s_hi(gg); ++j) n += nctw_stage(gg, j);
        return n;
    }
    // ---- loop-invariant twiddles kept in registers (PlanDesc::hoist) ------------------------------------------------
    // hreg[hoist_base(g, ord) + slot]: stage twiddle `slot` of the ord-th work item a thread runs in group g >= 1
    // (ord = iteration for ordinary groups; 2*iteration + {0, 1} for the items (u, m-u) of a paired group)
    KF_CE int hoist_ords(int g) const { return 2 * iters(g); }
    KF_CE int hoist_base(int g, int ord) const
    {
        int n = 0;
        for (int j = 1; j < g; ++j) n += hoist_ords(j) * nslots(j);
        return n + ord * nslots(g);
    }
    KF_CE int hoist_total() const { return G >= 2 ? hoist_base(G - 1, hoist_ords(G - 1)) : 0; }
    // is slot set (g, s, e) fetched from the table at run time?  (split mode derives the upper != 0 ones)
    KF_CE bool slot_fetched(int g, int s, int e, bool fixed) const
    {
        return digit(g, s, e) == 0 && (fixed || twmode != 1 || e / (W(g, s) * p[s]) == 0);
    }
    // register index of the first element whose upper digits (stages above s) form combination `up`
    KF_CE int upper_base(int g, int s, int up) const { return up * W(g, s) * p[s]; }
    // last group only (F_lo == 1): register that holds output k' + j*items(g)
    KF_CE int reg_of_j(int g, int j) const
    {
        for (int e = 0; e < R(g); ++e)
            if (kout(g, e) == j * items(g)) return e;
        return -1;
    }
    KF_CE int phys(int a) const { return logpad >= 31 ? a : a + (a >> logpad); }
    // phys(base + delta) == phys(base) + phys(delta) for every work item of group g?  (no carry out of the low
    // logpad bits).  Reads: base = kp*Flo*R + off (off < Flo), delta = e*Flo.  Writes: base = kp*Flo + off
    // (kp < m_hi), delta = kout(e)*Flo, a multiple of m_hi*Flo.
    KF_CE bool lin_rd(int g) const
    {
        if (logpad >= 31) return true;
        const int P = 1 << logpad, fl = Flo(g);
        return ((fl * R(g)) % P == 0) && (P % fl == 0 || fl % P == 0);
    }
    KF_CE bool lin_wr(int g) const
    {
        if (logpad >= 31) return true;
        const int P = 1 << logpad, u = mhi(g) * Flo(g);
        return (u % P == 0) || (P % u == 0);
    }
    // per-transform pitch of one exchange buffer (elements), made odd so that lanes which walk across
    // transforms (strided-input mapping) fall into different banks
    KF_CE int pitch() const { return phys(N - 1) + 1 + ((phys(N - 1) + 1) % 2 == 0 ? 1 : 0); }
    KF_CE int threads() const { return team * tpc; }
    KF_CE bool valid() const
    {
        int prod = 1, sum = 0;
        for (int s = 0; s < L; ++s) prod *= p[s];
        for (int g = 0; g < G; ++g) sum += glen[g];
        return prod == N && sum == L && L <= kMaxStages && G <= kMaxGroups && team > 0 && tpc > 0 && nslots(0) <= kMaxG0Slots && (twmode == 0 || nctw() <= kMaxCtw);
    }
};

}   // namespace kf
