// emul.cpp -- TEST-ONLY CPU emulator of the kernel bodies in kissfft_b200/csrc/kf_body.h.
//
// Purpose: exercise the index math of the kernels (autosort addressing, register-group digit bookkeeping,
// exchange-buffer parity, tile tails, real pre/post passes) in the container that has no GPU.  Every CUDA
// thread of a CTA becomes one std::thread and __syncthreads() becomes a std::barrier, so the very same
// template code that nvcc compiles for sm_100a runs here.  This file is never part of the product library and
// nothing in kissfft_b200/ calls it; GPU parity is established separately by the -m gpu tests.
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/kiss_fft.h"
#define KF_SCALAR_BYTES ((int)sizeof(kiss_fft_scalar))
#include "../../kissfft_b200/csrc/kf_twtab.h"
#include "../../kissfft_b200/csrc/kf_plan_list.h"

using namespace kf;
typedef Arith<kiss_fft_scalar> AT;
typedef AT::C CT;

struct EmuEnv {
    int tid_, nthreads_;
    long long bid_, nblocks_;
    unsigned char* smem_;
    std::barrier<>* bar_;
    int tid() const { return tid_; }
    int nthreads() const { return nthreads_; }
    long long bid() const { return bid_; }
    long long nblocks() const { return nblocks_; }
    void sync() const { bar_->arrive_and_wait(); }
    unsigned char* smem() const { return smem_; }
    // emulated warp shuffle: per-warp mailbox + per-warp barrier (all lanes of a warp call it convergently)
    struct WarpBox { unsigned long long slot[32]; std::barrier<>* bar; };
    WarpBox* boxes_;
    template <class V>
    V shfl(V v, int src_lane) const
    {
        WarpBox& wb = boxes_[tid_ / 32];
        unsigned long long raw = 0;
        std::memcpy(&raw, &v, sizeof(V));
        wb.slot[tid_ % 32] = raw;
        wb.bar->arrive_and_wait();
        raw = wb.slot[src_lane];
        wb.bar->arrive_and_wait();
        V r;
        std::memcpy(&r, &raw, sizeof(V));
        return r;
    }
    // emulated mbarrier: the 8 bytes hold a completion counter; the "bulk copy" is an immediate memcpy
    static std::atomic<unsigned long long>* ctr(void* bar) { return reinterpret_cast<std::atomic<unsigned long long>*>(bar); }
    void mbar_init(void* bar) const { ctr(bar)->store(0); }
    void mbar_fence_init() const {}
    void bulk_load(void* bar, void* dst, const void* src, unsigned bytes) const
    {
        if (bytes % 16 || ((uintptr_t)dst % 16) || ((uintptr_t)src % 16)) abort();   // cp.async.bulk alignment rules
        std::memcpy(dst, src, bytes);
        ctr(bar)->fetch_add(1, std::memory_order_release);
    }
    void mbar_expect(void*, unsigned) const {}
    // emulated tensor-map box: gather ncol adjacent elements of nrows rows (`row_stride` elements apart) starting at src
    void tensor_load_box(void* bar, void* dst, const void*, long long, int, long long, const void* src, long long row_stride, int ncol,
                         int nrows, int elem_bytes, bool last) const
    {
        if (((uintptr_t)dst % 128) || ((uintptr_t)src % 16) || ((size_t)row_stride * elem_bytes) % 16 || ((size_t)ncol * elem_bytes) % 16) abort();
        for (int r = 0; r < nrows; ++r)
            std::memcpy((char*)dst + (size_t)r * ncol * elem_bytes, (const char*)src + (size_t)r * row_stride * elem_bytes, (size_t)ncol * elem_bytes);
        if (last) ctr(bar)->fetch_add(1, std::memory_order_release);
    }
    void mbar_wait(void* bar, int k) const
    {
        while (ctr(bar)->load(std::memory_order_acquire) < (unsigned long long)k + 1) std::this_thread::yield();
    }
};

static KParams<AT> mk_params(int nfft, int inverse, const void* in, void* out, long long howmany, long long in_dist,
                             long long out_dist, long long in_stride, const void* tw, const void* stw)
{
    KParams<AT> P;
    P.ncols = 0;
    P.in_pdist = P.out_pdist = 0;
    P.npeers = 0;
    P.cols_per_peer = 0;
    P.peer_col_dist = 0;
    P.max_ctas = 0;
    P.in = (const CT*)in;
    P.out = (CT*)out;
    P.howmany = howmany;
    P.in_dist = in_dist;
    P.out_dist = out_dist;
    P.in_stride = in_stride;
    P.tw = (const CT*)tw;
    P.stw = (const CT*)stw;
    P.gtw = nullptr;
    const CT* h = (const CT*)tw;
    CT z{};
    P.pc.epi3 = AT::load((nfft % 3 == 0) ? h[nfft / 3] : z);
    P.pc.ya = AT::load((nfft % 5 == 0) ? h[nfft / 5] : z);
    P.pc.yb = AT::load((nfft % 5 == 0) ? h[2 * (nfft / 5)] : z);
    P.inverse = inverse;
    return P;
}

template <class F>
static void run_cta_grid(int nthreads, long long nblocks, size_t smem_bytes, F&& body)
{
    for (long long b = 0; b < nblocks; ++b) {
        std::vector<unsigned char> smem_store(smem_bytes + 256, 0xCD);
        unsigned char* smem_al = (unsigned char*)(((uintptr_t)smem_store.data() + 127) & ~(uintptr_t)127);
        std::barrier<> bar(nthreads);
        const int nwarps = (nthreads + 31) / 32;
        std::vector<EmuEnv::WarpBox> boxes(nwarps);
        std::vector<std::unique_ptr<std::barrier<>>> wbars;
        for (int w = 0; w < nwarps; ++w) {
            wbars.emplace_back(new std::barrier<>(std::min(32, nthreads - 32 * w)));
            boxes[w].bar = wbars.back().get();
        }
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                EmuEnv env{t, nthreads, b, nblocks, smem_al, &bar, boxes.data()};
                body(env);
            });
        for (auto& x : th) x.join();
    }
}

template <class PT, int MODE>
static int run_fused(KParams<AT>& P, long long nblocks)
{
    constexpr PlanDesc D = PT::D;
    std::vector<CT> gtw = build_gtw<AT, PT>(P.tw);
    P.gtw = gtw.data();
    fill_g0tw<AT, PT>(P, P.tw);
    // like the launcher: only the rows the alignment rules allow; the caller runs the rest on the run-time kernel
    const long long nfused = fused_rows<AT, PT, MODE>(P);
    P.howmany = nfused;
    if (nfused == 0) return 0;
    const size_t smem = FusedLayout<AT, PT, MODE>::kTotal;
    run_cta_grid(D.threads(), nblocks, smem, [&](EmuEnv& env) { fused_body<AT, PT, MODE>(P, env); });
    return (int)nfused;
}

typedef int (*fused_fn)(KParams<AT>&, long long);
struct Entry {
    int N;
    fused_fn fn[6];
};
#if defined(FIXED_POINT)
#define KF_4STEP(PT) nullptr, run_fused<PT, kC2CColCol>
#else
#define KF_4STEP(PT) run_fused<PT, kC2CColTw>, run_fused<PT, kC2CColCol>
#endif
#define KF_FUSED_ALL(PT) { PT::D.N, { run_fused<PT, kC2C>, run_fused<PT, kC2CCol>, run_fused<PT, kR2C>, run_fused<PT, kC2R>, KF_4STEP(PT) } }
#define KF_FUSED_C2C(PT) { PT::D.N, { run_fused<PT, kC2C>, nullptr, nullptr, nullptr, nullptr, nullptr } }
#define KF_FUSED_C2C_REAL(PT) { PT::D.N, { run_fused<PT, kC2C>, nullptr, run_fused<PT, kR2C>, run_fused<PT, kC2R>, nullptr, nullptr } }
#define KF_FUSED_C2C_COL(PT) { PT::D.N, { run_fused<PT, kC2C>, run_fused<PT, kC2CCol>, nullptr, nullptr, KF_4STEP(PT) } }
#define KF_FUSED_COL(PT) { PT::D.N, { nullptr, run_fused<PT, kC2CCol>, nullptr, nullptr, KF_4STEP(PT) } }
#define KF_FUSED_R2C(PT) { PT::D.N, { nullptr, nullptr, run_fused<PT, kR2C>, nullptr, nullptr, nullptr } }
#define KF_FUSED_C2R(PT) { PT::D.N, { nullptr, nullptr, nullptr, run_fused<PT, kC2R>, nullptr, nullptr } }
#define KF_FUSED_REAL(PT) { PT::D.N, { nullptr, nullptr, run_fused<PT, kR2C>, run_fused<PT, kC2R>, nullptr, nullptr } }
#define KF_ROW(tag, modes) KF_FUSED_##modes(tag),
static const Entry kTable[] = { KF_PLAN_LIST(KF_ROW) };

// ---- experimental plans: variants that are NOT in the product's plan list (candidates for the next tuning round), so
// that their index math can be validated here before any GPU time is spent on them
#include "experimental_plans.h"
static const Entry kExperimental[] = { KF_EXPERIMENTAL_LIST(KF_ROW) { 0, { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr } } };
extern "C" int emul_num_experimental(void) { return (int)(sizeof(kExperimental) / sizeof(kExperimental[0])) - 1; }
extern "C" int emul_experimental_nfft(int i) { return kExperimental[i].N; }
extern "C" int emul_experimental_has_mode(int i, int mode) { return kExperimental[i].fn[mode] != nullptr; }
extern "C" int emul_fused_experimental(int idx, int mode, int inverse, const void* in, void* out, long long howmany, long long in_dist,
                                       long long out_dist, long long in_stride, const void* tw, const void* stw, long long nblocks)
{
    const Entry& e = kExperimental[idx];
    if (!e.fn[mode]) return -1;
    KParams<AT> P = mk_params(e.N, inverse, in, out, howmany, in_dist, out_dist, in_stride, tw, stw);
    return e.fn[mode](P, nblocks);
}

extern "C" int emul_num_plans(void) { return (int)(sizeof(kTable) / sizeof(kTable[0])); }
extern "C" int emul_plan_nfft(int i) { return kTable[i].N; }
// first table entry that serves (N, mode) wins, exactly like find_fused in kf_launch.cu
extern "C" int emul_plan_has_mode(int i, int mode)
{
    if (!kTable[i].fn[mode]) return 0;
    for (int j = 0; j < i; ++j)
        if (kTable[j].N == kTable[i].N && kTable[j].fn[mode]) return 0;
    return 1;
}

// returns the number of leading rows processed (the launcher sends the rest to the run-time kernel), or -1 when no
// fused plan is registered for (nfft, mode)
extern "C" int emul_fused(int nfft, int mode, int inverse, const void* in, void* out, long long howmany, long long in_dist,
                          long long out_dist, long long in_stride, const void* tw, const void* stw, long long nblocks)
{
    for (const Entry& e : kTable)
        if (e.N == nfft && e.fn[mode]) {
            KParams<AT> P = mk_params(nfft, inverse, in, out, howmany, in_dist, out_dist, in_stride, tw, stw);
            return e.fn[mode](P, nblocks);
        }
    return -1;
}

// one pass of the four-step transform, parameters exactly as kf_launch.cu:kfcu_exec_fourstep builds them
extern "C" int emul_fourstep(int nfft, int step, int inverse, const void* in, void* out, long long nrows, long long ncols,
                             const void* tw, const void* twbig, long long nblocks)
{
    const int mode = step == 0 ? (int)kC2CColTw : (int)kC2CColCol;
    for (const Entry& e : kTable)
        if (e.N == nfft && e.fn[mode]) {
            const long long N = (long long)nfft * ncols;
            KParams<AT> P = mk_params(nfft, inverse, in, out, nrows * ncols, 1, step == 0 ? nfft : 1, ncols, tw, step == 0 ? twbig : nullptr);
            P.ncols = ncols;
            P.in_pdist = N;
            P.out_pdist = N;
            return e.fn[mode](P, nblocks);
        }
    return -1;
}

// column-ring variants (tensor-map TMA input ring, KF_COLRING_LIST): mode kC2CCol (transposing, out row = column),
// kC2CColTw / kC2CColCol with the geometry kf_launch.cu builds.  Returns -1 when this datatype build has no ring
// plan for nfft, -2 when the geometry is not eligible (the launcher would then take the direct-load plan).
#if defined(FIXED_POINT)
#define KF_RING_TW(PT, on) nullptr
#else
#define KF_RING_TW(PT, on) ((on) ? run_fused<PT, kC2CColTw> : nullptr)
#endif
#define KF_RING_ROW(PT, c, tw, cc) { PT::D.N, { nullptr, (c) ? run_fused<PT, kC2CCol> : nullptr, nullptr, nullptr, KF_RING_TW(PT, tw), \
                                     (cc) ? run_fused<PT, kC2CColCol> : nullptr } },
static const Entry kRingTable[] = { KF_COLRING_LIST(KF_RING_ROW) { 0, { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr } } };
extern "C" int emul_colring(int nfft, int mode, int inverse, const void* in, void* out, long long nplanes, long long ncols,
                            long long col_stride, long long in_pdist, long long out_pdist, long long out_dist, const void* tw,
                            const void* twbig, long long nblocks)
{
    for (const Entry& e : kRingTable)
        if (e.N == nfft && e.fn[mode]) {
            KParams<AT> P = mk_params(nfft, inverse, in, out, nplanes * ncols, 1, out_dist, col_stride, tw, twbig);
            P.ncols = ncols;
            P.in_pdist = in_pdist;
            P.out_pdist = out_pdist;
            if (ncols % 8 != 0) return -2;
            return e.fn[mode](P, nblocks);
        }
    return -1;
}

extern "C" int emul_generic(int nfft, int mode, int inverse, const int* factors, int nstages, const void* in, void* out,
                            long long howmany, long long in_dist, long long out_dist, long long in_stride, const void* tw,
                            const void* stw, int tpc, int nthreads, long long nblocks)
{
    GParams<AT> G;
    G.k = mk_params(nfft, inverse, in, out, howmany, in_dist, out_dist, in_stride, tw, stw);
    G.plan.N = nfft;
    G.plan.L = nstages;
    for (int s = 0; s < nstages; ++s) { G.plan.p[s] = factors[2 * s]; G.plan.m[s] = factors[2 * s + 1]; }
    G.mode = mode;
    G.tpc = tpc;
    const size_t smem = (size_t)2 * tpc * nfft * sizeof(CT);
    run_cta_grid(nthreads, nblocks, smem, [&](EmuEnv& env) { generic_body<AT>(G, env); });
    return 0;
}

// ---- multi-pass path: one stage / one real split pass per call ------------------------------------------------
extern "C" int emul_stage(int nfft, int inverse, int p, int m, int F, const void* in, void* out, long long batch,
                          long long in_dist, long long out_dist, long long in_stride, int first, int last, const void* tw,
                          int nthreads, long long nblocks)
{
    KParams<AT> P = mk_params(nfft, inverse, in, out, batch, in_dist, out_dist, in_stride, tw, nullptr);
    StageParams<AT> S;
    S.in = (const CT*)in; S.out = (CT*)out; S.batch = batch;
    S.in_dist = in_dist; S.out_dist = out_dist; S.in_stride = in_stride;
    S.N = nfft; S.p = p; S.m = m; S.F = F; S.first = first; S.last = last;
    S.tw = (const CT*)tw; S.pc = P.pc; S.inverse = inverse;
    run_cta_grid(nthreads, nblocks, 16, [&](EmuEnv& env) { stage_body<AT>(S, env); });
    return 0;
}

extern "C" int emul_realpass(int nc, int post, const void* in, void* out, long long batch, long long in_dist, long long out_dist,
                             const void* stw, int nthreads, long long nblocks)
{
    RealPassParams<AT> S;
    S.in = (const CT*)in; S.out = (CT*)out; S.batch = batch; S.in_dist = in_dist; S.out_dist = out_dist;
    S.nc = nc; S.post = post; S.stw = (const CT*)stw;
    run_cta_grid(nthreads, nblocks, 16, [&](EmuEnv& env) { realpass_body<AT>(S, env); });
    return 0;
}

// ---- fused fast convolution (float / double) -------------------------------------------------------------------------
#if !defined(FIXED_POINT)
template <class PT>
static int run_fastconv(const void* in, void* out, long long nblocks, long long ngood, const void* h, const void* tw_f,
                        const void* tw_i, long long grid)
{
    constexpr PlanDesc D = PT::D;
    std::vector<CT> gf = build_gtw<AT, PT>((const CT*)tw_f), gi = build_gtw<AT, PT>((const CT*)tw_i);
    FCParams<AT> P;
    P.in = (const CT*)in; P.out = (CT*)out; P.nblocks = nblocks; P.ngood = ngood; P.h = (const CT*)h;
    fill_fc_side<AT, PT>(P.fwd, (const CT*)tw_f, (const CT*)tw_f, gf.data());
    fill_fc_side<AT, PT>(P.inv, (const CT*)tw_i, (const CT*)tw_i, gi.data());
    const size_t smem = (D.G >= 2) ? (size_t)2 * D.tpc * D.pitch() * sizeof(CT) : 16;
    run_cta_grid(D.threads(), grid, smem, [&](EmuEnv& env) { fastconv_body<AT, PT>(P, env); });
    return 0;
}
typedef int (*fc_fn)(const void*, void*, long long, long long, const void*, const void*, const void*, long long);
struct FCEntry { int N; fc_fn fn; };
#define KF_FC_ROW(tag) { tag::D.N, run_fastconv<tag> },
static const FCEntry kFCTable[] = { KF_FASTCONV_LIST(KF_FC_ROW) };
extern "C" int emul_fastconv(int nfft, const void* in, void* out, long long nblocks, long long ngood, const void* h,
                             const void* tw_f, const void* tw_i, long long grid)
{
    for (const FCEntry& e : kFCTable)
        if (e.N == nfft) return e.fn(in, out, nblocks, ngood, h, tw_f, tw_i, grid);
    return -1;
}
#endif
