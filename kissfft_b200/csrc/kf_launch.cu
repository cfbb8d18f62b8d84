// kf_launch.cu -- kernel selection and launch (the only translation unit that instantiates kernels).
//
// Built once per datatype (FIXED_POINT / kiss_fft_scalar macros, see include/kiss_fft.h).  The host layer
// (kf_api.c) hands over a kfcu_plan that carries the reference's radix schedule and twiddle tables; this file
// picks the compile-time fused plan when one is registered for the length (kf_plans.inc) and otherwise the
// run-time shared-memory kernel.  There is no CPU path: if no kernel applies the call fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/kiss_fft.h"
#include "kf_internal.h"
#include "kf_kernels.cuh"
#include "kf_twtab.h"
#include "kf_tmap.h"
#define KF_SCALAR_BYTES ((int)sizeof(kiss_fft_scalar))
#include "kf_plan_list.h"

using namespace kf;

typedef Arith<kiss_fft_scalar> AT;
typedef AT::C CT;
static_assert(sizeof(CT) == sizeof(kiss_fft_cpx), "storage complex must match kiss_fft_cpx");

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_force_generic{0};
static thread_local int g_grid_limit = 0;  // launches of THIS host thread: 0 = fill the device; > 0 caps the persistent grid
static thread_local int g_sm_reserve = 0;  // launches of THIS host thread leave that many SMs free (a concurrent kernel owns them)

struct DeviceInfo {
    int sms = 0;
    int max_smem_optin = 0;
};
static DeviceInfo device_info()
{
    static DeviceInfo cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cache[dev].sms == 0) {
        cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&cache[dev].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return cache[dev];
}

static KParams<AT> make_params(const kfcu_plan* pl, const void* d_in, void* d_out, long long howmany, long long in_dist,
                               long long out_dist, long long in_stride)
{
    KParams<AT> P;
    memset(P.tmap, 0, sizeof(P.tmap));
    P.ncols = 0;
    P.in_pdist = P.out_pdist = 0;
    P.npeers = 0;
    P.cols_per_peer = 0;
    P.peer_col_dist = 0;
    P.max_ctas = 0;
    P.in = (const CT*)d_in;
    P.out = (CT*)d_out;
    P.howmany = howmany;
    P.in_dist = in_dist;
    P.out_dist = out_dist;
    P.in_stride = in_stride;
    P.tw = (const CT*)pl->d_tw;
    P.stw = (const CT*)pl->d_stw;
    P.gtw = nullptr;
    const CT* h = (const CT*)pl->h_tw;
    const int N = pl->nfft;
    // constants of kf_bfly3 / kf_bfly5 (kiss_fft.c:99, 143-144): twiddles[fstride*m] with fstride*m == N/p
    CT z{};
    P.pc.epi3 = AT::load((N % 3 == 0) ? h[N / 3] : z);
    P.pc.ya = AT::load((N % 5 == 0) ? h[N / 5] : z);
    P.pc.yb = AT::load((N % 5 == 0) ? h[2 * (N / 5)] : z);
    P.inverse = pl->inverse;
    return P;
}

// ---- fused plans ------------------------------------------------------------------------------------------
typedef int (*fused_launch_fn)(kfcu_plan*, KParams<AT>&, cudaStream_t);

static std::mutex g_gtw_mutex;
static int launch_generic(int mode, const kfcu_plan* pl, const KParams<AT>& P, cudaStream_t st);

// rows [first, first+count) of a call
static KParams<AT> sub_rows(const KParams<AT>& P, long long first, long long count)
{
    KParams<AT> Q = P;
    Q.in = P.in + first * P.in_dist;
    Q.out = P.out + first * P.out_dist;
    Q.howmany = count;
    return Q;
}

template <class PT, int MODE>
static int launch_fused(kfcu_plan* pl, KParams<AT>& P, cudaStream_t st)
{
    constexpr PlanDesc D = PT::D;
    // stage-twiddle tables of this plan: built once per (device plan) from the host twiddles, then cached
    constexpr int kSlot = (MODE == kC2CColTw || MODE == kC2CColCol) ? (int)kC2CCol : MODE;   // same plan tag as the column mode
    if (D.gtw_total() > 0 && !pl->d_gtw[kSlot]) {
        std::lock_guard<std::mutex> lk(g_gtw_mutex);
        if (!pl->d_gtw[kSlot]) {
            std::vector<CT> tab = build_gtw<AT, PT>((const CT*)pl->h_tw);
            void* d = nullptr;
            cudaError_t e = cudaMalloc(&d, tab.size() * sizeof(CT));
            if (e != cudaSuccess) return (int)e;
            e = cudaMemcpy(d, tab.data(), tab.size() * sizeof(CT), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); return (int)e; }
            pl->d_gtw[kSlot] = d;
        }
    }
    P.gtw = (const CT*)pl->d_gtw[kSlot];
    fill_g0tw<AT, PT>(P, (const CT*)pl->h_tw);
    // rows the fused kernel may take (alignment rules of the bulk-async ring); the remainder, normally none,
    // goes to the run-time kernel
    const long long nfused = fused_rows<AT, PT, MODE>(P);
    if (nfused < P.howmany) {
        int rc = launch_generic(MODE, pl, sub_rows(P, nfused, P.howmany - nfused), st);
        if (rc != 0 || nfused == 0) return rc;
        P = sub_rows(P, 0, nfused);
    }
    if constexpr (FusedLayout<AT, PT, MODE>::kColRing) {
        // tensor-map input ring: the caller (launch_col) has checked col_ring_ok for this plan
        const int rc = col_ring_encode<AT, PT, MODE>(P);
        if (rc != 0) return rc;
    }
    auto kern = kf_fused_kernel<AT, PT, MODE>;
    constexpr size_t smem = FusedLayout<AT, PT, MODE>::kTotal;
    static_assert(smem <= 232448, "plan exceeds the 227 KiB of shared memory a CTA can opt in to");
    static int blocks_per_sm[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (blocks_per_sm[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, D.threads(), smem);
        if (e != cudaSuccess) return (int)e;
        if (nb < 1) return (int)cudaErrorLaunchOutOfResources;
        blocks_per_sm[dev] = nb;
    }
    const long long ntiles = (P.howmany + D.tpc - 1) / D.tpc;
    int sms = device_info().sms;
    if (g_sm_reserve > 0) sms = sms - g_sm_reserve > 8 ? sms - g_sm_reserve : 8;
    long long grid = (long long)sms * blocks_per_sm[dev];
    if (grid > ntiles) grid = ntiles;
    if (const int lim = g_grid_limit; lim > 0 && grid > lim) grid = lim;
    if (P.max_ctas > 0 && grid > P.max_ctas) grid = P.max_ctas;      // per-launch cap (link-bound launches share the SMs)
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, D.threads(), smem, st>>>(P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

struct FusedEntry {
    int N;
    fused_launch_fn fn[6];   // indexed by Mode; null = not instantiated
};

// the two four-step modes ride on every plan that serves the column mode (float / double only)
#if defined(FIXED_POINT)
#define KF_4STEP(PT) nullptr, launch_fused<PT, kC2CColCol>
#else
#define KF_4STEP(PT) launch_fused<PT, kC2CColTw>, launch_fused<PT, kC2CColCol>
#endif
#define KF_FUSED_ALL(PT) { PT::D.N, { launch_fused<PT, kC2C>, launch_fused<PT, kC2CCol>, launch_fused<PT, kR2C>, launch_fused<PT, kC2R>, KF_4STEP(PT) } }
#define KF_FUSED_C2C(PT) { PT::D.N, { launch_fused<PT, kC2C>, nullptr, nullptr, nullptr, nullptr, nullptr } }
#define KF_FUSED_C2C_REAL(PT) { PT::D.N, { launch_fused<PT, kC2C>, nullptr, launch_fused<PT, kR2C>, launch_fused<PT, kC2R>, nullptr, nullptr } }
#define KF_FUSED_C2C_COL(PT) { PT::D.N, { launch_fused<PT, kC2C>, launch_fused<PT, kC2CCol>, nullptr, nullptr, KF_4STEP(PT) } }
#define KF_FUSED_COL(PT) { PT::D.N, { nullptr, launch_fused<PT, kC2CCol>, nullptr, nullptr, KF_4STEP(PT) } }
#define KF_FUSED_R2C(PT) { PT::D.N, { nullptr, nullptr, launch_fused<PT, kR2C>, nullptr, nullptr, nullptr } }
#define KF_FUSED_C2R(PT) { PT::D.N, { nullptr, nullptr, nullptr, launch_fused<PT, kC2R>, nullptr, nullptr } }
#define KF_FUSED_REAL(PT) { PT::D.N, { nullptr, nullptr, launch_fused<PT, kR2C>, launch_fused<PT, kC2R>, nullptr, nullptr } }

#include "kf_plans.inc"

// column-ring variants (tensor-map TMA input) of the column plans: preferred whenever the call's geometry allows
struct ColRingEntry {
    int N;
    fused_launch_fn fn[6];
    bool (*ok)(const KParams<AT>&);
};
// X(tag, serves kC2CCol, serves kC2CColTw, serves kC2CColCol)
#if defined(FIXED_POINT)
#define KF_RING_TW(PT, on) nullptr
#else
#define KF_RING_TW(PT, on) ((on) ? launch_fused<PT, kC2CColTw> : nullptr)
#endif
#define KF_COLRING_ROW(PT, c, tw, cc) { PT::D.N, { nullptr, (c) ? launch_fused<PT, kC2CCol> : nullptr, nullptr, nullptr, KF_RING_TW(PT, tw), \
                                        (cc) ? launch_fused<PT, kC2CColCol> : nullptr }, col_ring_ok<AT, PT> },
static const ColRingEntry kColRingTable[] = { KF_COLRING_LIST(KF_COLRING_ROW) { 0, { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }, nullptr } };
static std::atomic<int> g_no_colring{0};

static const FusedEntry* find_fused(int nfft, int mode);
static const char* env_cached(const char* name)
{
    return getenv(name);
}

// column-mode launch: ring variant when eligible, else the direct-load plan, else the run-time kernel (kC2CCol only)
static int launch_col(int mode, kfcu_plan* plan, KParams<AT>& P, cudaStream_t st)
{
    static const bool off = [] { const char* e = env_cached("KISSFFT_COL_TMA"); return e && e[0] == '0'; }();
    if (!off && !g_force_generic.load() && !g_no_colring.load())
        for (const ColRingEntry& e : kColRingTable)
            if (e.N == plan->nfft && e.fn[mode] && e.ok(P)) return e.fn[mode](plan, P, st);
    if (const FusedEntry* fe = find_fused(plan->nfft, mode)) return fe->fn[mode](plan, P, st);
    if (mode == kC2CCol) return launch_generic(kC2CCol, plan, P, st);
    return KFCU_EINVAL;
}

static const FusedEntry* find_fused(int nfft, int mode)
{
    if (g_force_generic.load()) return nullptr;
    for (const FusedEntry& e : kFusedTable)
        if (e.N == nfft && e.fn[mode]) return &e;
    return nullptr;
}

// ---- generic kernel ---------------------------------------------------------------------------------------
static int launch_generic(int mode, const kfcu_plan* pl, const KParams<AT>& P, cudaStream_t st)
{
    const DeviceInfo di = device_info();
    const size_t per = (size_t)2 * pl->nfft * sizeof(CT);
    if (per > (size_t)di.max_smem_optin) return KFCU_ETOOBIG;
    GParams<AT> G;
    G.k = P;
    G.plan.N = pl->nfft;
    G.plan.L = pl->nstages;
    for (int s = 0; s < pl->nstages; ++s) { G.plan.p[s] = pl->p[s]; G.plan.m[s] = pl->m[s]; }
    G.mode = mode;
    // transforms per CTA: fill ~32 KiB of exchange buffers; column mode wants >= 64-byte row segments
    int tpc = (int)((32 * 1024) / per);
    if (mode == kC2CCol) { int want = (int)(64 / sizeof(CT)); if (want < 1) want = 1; if (tpc < want) tpc = want; }
    if (tpc < 1) tpc = 1;
    while (tpc > 1 && (size_t)tpc * per > (size_t)di.max_smem_optin) --tpc;
    if ((long long)tpc > P.howmany) tpc = (int)P.howmany;
    G.tpc = tpc;
    const size_t smem = (size_t)tpc * per;
    auto kern = kf_generic_kernel<AT>;
    static int attr_set[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev] = 1;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem);
    if (e != cudaSuccess) return (int)e;
    if (nb < 1) return (int)cudaErrorLaunchOutOfResources;
    const long long ntiles = (P.howmany + tpc - 1) / tpc;
    long long grid = (long long)di.sms * nb;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, 256, smem, st>>>(G);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// ---- C interface ------------------------------------------------------------------------------------------
extern "C" int kfcu_exec(int mode, kfcu_plan* plan, const void* d_in, void* d_out, long long howmany,
                         long long in_dist, long long out_dist, long long in_stride, void* stream)
{
    if (!plan || !d_in || !d_out || mode < 0 || mode > 3 || howmany < 0) return KFCU_EINVAL;
    if (howmany == 0) return 0;
    if ((mode == kR2C || mode == kC2R) && !plan->d_stw && plan->nfft > 1) return KFCU_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    KParams<AT> P = make_params(plan, d_in, d_out, howmany, in_dist, out_dist, in_stride);
    if (mode == kC2CCol) return launch_col(mode, plan, P, st);
    if (const FusedEntry* fe = find_fused(plan->nfft, mode)) return fe->fn[mode](plan, P, st);
    return launch_generic(mode, plan, P, st);
}

// plane-batched column pass: plane p, column c: in[p*in_pdist + c + j*col_stride] (j < nfft) -> out[p*out_pdist + c*nfft + k]
extern "C" int kfcu_exec_planes(kfcu_plan* plan, const void* d_in, void* d_out, long long nplanes, long long ncols,
                                long long col_stride, long long in_pdist, long long out_pdist, void* stream)
{
    if (!plan || !d_in || !d_out || nplanes < 0 || ncols < 0) return KFCU_EINVAL;
    if (nplanes == 0 || ncols == 0) return 0;
    KParams<AT> P = make_params(plan, d_in, d_out, nplanes * ncols, 1, plan->nfft, col_stride);
    P.ncols = ncols;
    P.in_pdist = in_pdist;
    P.out_pdist = out_pdist;
    return launch_col(kC2CCol, plan, P, (cudaStream_t)stream);
}

// the two passes of the four-step transform of rows of length N = N1 * N2 (float / double; kf_api.c:kf_exec_fourstep).
//   step 0 (kC2CColTw): plan of length N1; plane = one row viewed as [N1][N2]; column n2 -> out[p*N + n2*N1 + k1] * W_N^(n2*k1)
//   step 1 (kC2CColCol): plan of length N2; plane = the intermediate array viewed as [N2][N1]; column k1 -> out[p*N + k2*N1 + k1]
// d_twbig: the N twiddles of the long transform (step 0 only).  Returns KFCU_EINVAL when no fused column plan serves nfft.
extern "C" int kfcu_exec_fourstep(kfcu_plan* plan, int step, const void* d_in, void* d_out, long long nrows, long long ncols,
                                  const void* d_twbig, void* stream)
{
    if (!plan || !d_in || !d_out || nrows < 0 || ncols < 1 || step < 0 || step > 1) return KFCU_EINVAL;
    if (nrows == 0) return 0;
    const int mode = step == 0 ? (int)kC2CColTw : (int)kC2CColCol;
    if (!find_fused(plan->nfft, mode)) return KFCU_EINVAL;
    const long long N = (long long)plan->nfft * ncols;
    KParams<AT> P = make_params(plan, d_in, d_out, nrows * ncols, 1, step == 0 ? plan->nfft : 1, ncols);
    P.ncols = ncols;
    P.in_pdist = N;
    P.out_pdist = N;
    if (step == 0) {
        if (!d_twbig) return KFCU_EINVAL;
        P.stw = (const CT*)d_twbig;
    }
    return launch_col(mode, plan, P, (cudaStream_t)stream);
}

extern "C" int kfcu_has_colcol(int nfft) { return find_fused(nfft, kC2CColCol) != nullptr; }

// does a tensor-map input-ring variant serve the transposing column pass of this length?
extern "C" int kfcu_has_colring(int nfft)
{
    for (const ColRingEntry& e : kColRingTable)
        if (e.N == nfft && e.fn[kC2CCol]) return 1;
    return 0;
}

extern "C" int kfcu_has_fourstep(int nfft)
{
    return find_fused(nfft, kC2CColTw) != nullptr && find_fused(nfft, kC2CColCol) != nullptr;
}

// the same pass with the columns of every plane split into npeers blocks; block s goes through peers[s]
extern "C" int kfcu_exec_planes_peers(kfcu_plan* plan, const void* d_in, void* const* peers, int npeers, long long nplanes,
                                      long long cols_per_peer, long long peer_col_dist, long long col_stride, long long in_pdist,
                                      long long out_pdist, long long out_col_dist, int max_ctas, void* stream)
{
    if (!plan || !d_in || !peers || npeers < 1 || npeers > 16 || nplanes < 0 || cols_per_peer < 1 || peer_col_dist < cols_per_peer) return KFCU_EINVAL;
    if (nplanes == 0) return 0;
    const long long ncols = cols_per_peer * npeers;
    // row (plane p, column c of peer s) lands at peers[s] + p*out_pdist + c*out_col_dist (default: rows back to back)
    KParams<AT> P = make_params(plan, d_in, peers[0], nplanes * ncols, 1, out_col_dist > 0 ? out_col_dist : plan->nfft, col_stride);
    P.ncols = ncols;
    P.in_pdist = in_pdist;
    P.out_pdist = out_pdist;
    P.npeers = npeers;
    P.cols_per_peer = cols_per_peer;
    P.peer_col_dist = peer_col_dist;
    for (int s = 0; s < npeers; ++s) P.peer[s] = (CT*)peers[s];
    P.max_ctas = max_ctas;
    return launch_col(kC2CCol, plan, P, (cudaStream_t)stream);
}

// ---- fused fast convolution (float / double) ------------------------------------------------------------------
#if !defined(FIXED_POINT)
typedef int (*fc_launch_fn)(kfcu_plan*, kfcu_plan*, const void*, void*, long long, long long, const void*, cudaStream_t);

template <class PT>
static int ensure_gtw(kfcu_plan* pl, int slot)
{
    constexpr PlanDesc D = PT::D;
    if (D.gtw_total() > 0 && !pl->d_gtw[slot]) {
        std::lock_guard<std::mutex> lk(g_gtw_mutex);
        if (!pl->d_gtw[slot]) {
            std::vector<CT> tab = build_gtw<AT, PT>((const CT*)pl->h_tw);
            void* d = nullptr;
            cudaError_t e = cudaMalloc(&d, tab.size() * sizeof(CT));
            if (e != cudaSuccess) return (int)e;
            e = cudaMemcpy(d, tab.data(), tab.size() * sizeof(CT), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); return (int)e; }
            pl->d_gtw[slot] = d;
        }
    }
    return 0;
}

template <class PT>
static int launch_fastconv(kfcu_plan* fwd, kfcu_plan* inv, const void* d_in, void* d_out, long long nblocks, long long ngood,
                           const void* d_h, cudaStream_t st)
{
    constexpr PlanDesc D = PT::D;
    auto kern = kf_fastconv_kernel<AT, PT>;
    constexpr size_t smem = (D.G >= 2) ? (size_t)2 * D.tpc * D.pitch() * sizeof(CT) : 0;
    static_assert(smem <= 232448, "fast-convolution plan exceeds shared memory");
    // the fast-convolution plan has its own slot (4) in the per-plan table cache
    int rc = ensure_gtw<PT>(fwd, 4);
    if (!rc) rc = ensure_gtw<PT>(inv, 4);
    if (rc) return rc;
    FCParams<AT> P;
    P.in = (const CT*)d_in;
    P.out = (CT*)d_out;
    P.nblocks = nblocks;
    P.ngood = ngood;
    P.h = (const CT*)d_h;
    fill_fc_side<AT, PT>(P.fwd, (const CT*)fwd->h_tw, (const CT*)fwd->d_tw, (const CT*)fwd->d_gtw[4]);
    fill_fc_side<AT, PT>(P.inv, (const CT*)inv->h_tw, (const CT*)inv->d_tw, (const CT*)inv->d_gtw[4]);
    static int blocks_per_sm[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (blocks_per_sm[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, D.threads(), smem);
        if (e != cudaSuccess) return (int)e;
        if (nb < 1) return (int)cudaErrorLaunchOutOfResources;
        blocks_per_sm[dev] = nb;
    }
    const long long ntiles = (nblocks + D.tpc - 1) / D.tpc;
    long long grid = (long long)device_info().sms * blocks_per_sm[dev];
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, D.threads(), smem, st>>>(P);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

struct FCEntry { int N; fc_launch_fn fn; };
#define KF_FC_ROW(tag) { tag::D.N, launch_fastconv<tag> },
static const FCEntry kFCTable[] = { KF_FASTCONV_LIST(KF_FC_ROW) };
#endif

extern "C" int kfcu_has_fastconv(int nfft)
{
#if !defined(FIXED_POINT)
    for (const FCEntry& e : kFCTable)
        if (e.N == nfft) return 1;
#endif
    (void)nfft;
    return 0;
}

extern "C" int kfcu_fastconv(kfcu_plan* fwd, kfcu_plan* inv, const void* d_in, void* d_out, long long nblocks, long long ngood,
                             const void* d_h, void* stream)
{
#if !defined(FIXED_POINT)
    if (!fwd || !inv || !d_in || !d_out || !d_h || fwd->nfft != inv->nfft || ngood < 1 || ngood > fwd->nfft) return KFCU_EINVAL;
    if (nblocks <= 0) return 0;
    for (const FCEntry& e : kFCTable)
        if (e.N == fwd->nfft) return e.fn(fwd, inv, d_in, d_out, nblocks, ngood, d_h, (cudaStream_t)stream);
#endif
    return KFCU_ETOOBIG;
}

// ---- flags between the GPUs of the slab transform (kf_mgpu.c): "block (chunk, source rank) has landed" -------------------
// flags[r] = rank r's flag words, mapped into this process; word (slot * 16 + source rank) of the receiver is set to the
// epoch by the source once its stores to that receiver are complete (the kernel that made them precedes this one on the
// stream; the system-scope fence orders them before the flag).
__global__ void kf_peer_signal_kernel(unsigned* const* flags, int nranks, int rank, int slot, unsigned epoch)
{
    const int s = (int)threadIdx.x;
    if (s < nranks) {
        __threadfence_system();
        volatile unsigned* f = flags[s] + slot * 16 + rank;
        *f = epoch;
        __threadfence_system();
    }
}
// waits until every source rank has signalled `slot` for this epoch (epochs only grow; wrap-around safe comparison).
// Gives up after ~20 s and traps, so a lost peer cannot hang the GPU.
__global__ void kf_peer_wait_kernel(const unsigned* mine, int nranks, int slot, unsigned epoch)
{
    const int s = (int)threadIdx.x;
    if (s < nranks) {
        const volatile unsigned* f = mine + slot * 16 + s;
        const long long t0 = clock64();
        while ((int)(*f - epoch) < 0) {
            if (clock64() - t0 > 40000000000LL) __trap();
        }
        __threadfence_system();
    }
}

extern "C" int kfcu_peer_signal(void* const* d_flag_ptrs, int nranks, int rank, int slot, unsigned epoch, void* stream)
{
    if (!d_flag_ptrs || nranks < 1 || nranks > 16) return KFCU_EINVAL;
    kf_peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned* const*)d_flag_ptrs, nranks, rank, slot, epoch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_peer_wait(const void* d_my_flags, int nranks, int slot, unsigned epoch, void* stream)
{
    if (!d_my_flags || nranks < 1 || nranks > 16) return KFCU_EINVAL;
    kf_peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned*)d_my_flags, nranks, slot, epoch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// ---- pieces of the unfused fast convolution (any nfft; kf_api.c:kf_fastconv_unfused) ------------------------------------
extern "C" int kfcu_gather_blocks(const void* d_in, void* d_out, long long nblocks, int len, long long advance, int is_real, void* stream)
{
    if (!d_in || !d_out || nblocks < 0 || len < 1 || advance < 1) return KFCU_EINVAL;
    if (nblocks == 0) return 0;
    long long grid = (nblocks * len + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (grid > cap) grid = cap;
    if (is_real)
        kf_gather_blocks_kernel<kiss_fft_scalar><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const kiss_fft_scalar*)d_in, (kiss_fft_scalar*)d_out, nblocks, len, advance);
    else
        kf_gather_blocks_kernel<CT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const CT*)d_in, (CT*)d_out, nblocks, len, advance);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_cmul_rows(void* d_x, const void* d_h, long long rows, int n, void* stream)
{
#if !defined(FIXED_POINT)
    if (!d_x || !d_h || rows < 0 || n < 1) return KFCU_EINVAL;
    if (rows == 0) return 0;
    long long grid = (rows * n + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (grid > cap) grid = cap;
    kf_cmul_rows_kernel<AT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((CT*)d_x, (const CT*)d_h, rows, n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
#else
    (void)d_x; (void)d_h; (void)rows; (void)n; (void)stream;
    return KFCU_EINVAL;
#endif
}

// ---- multi-pass path (lengths beyond the shared-memory kernels) -------------------------------------------------
extern "C" int kfcu_stage(const kfcu_plan* plan, int s, const void* d_in, void* d_out, long long batch, long long in_dist,
                          long long out_dist, long long in_stride, int first, int last, void* stream)
{
    if (!plan || !d_in || !d_out || s < 0 || s >= plan->nstages) return KFCU_EINVAL;
    if (batch <= 0) return 0;
    KParams<AT> P = make_params(plan, d_in, d_out, batch, in_dist, out_dist, in_stride);
    StageParams<AT> S;
    S.in = (const CT*)d_in;
    S.out = (CT*)d_out;
    S.batch = batch;
    S.in_dist = in_dist;
    S.out_dist = out_dist;
    S.in_stride = in_stride;
    S.N = plan->nfft;
    S.p = plan->p[s];
    S.m = plan->m[s];
    int F = 1;
    for (int j = 0; j < s; ++j) F *= plan->p[j];
    S.F = F;
    S.first = first;
    S.last = last;
    S.tw = P.tw;
    S.pc = P.pc;
    S.inverse = plan->inverse;
    const bool small = (S.p == 2 || S.p == 3 || S.p == 4 || S.p == 5);
    const long long work = batch * (small ? plan->nfft / S.p : plan->nfft);
    long long grid = (work + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (grid > cap) grid = cap;
    kf_stage_kernel<AT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(S);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_realpass(const kfcu_plan* plan, int post, const void* d_in, void* d_out, long long batch, long long in_dist,
                             long long out_dist, void* stream)
{
    if (!plan || !d_in || !d_out || (!plan->d_stw && plan->nfft > 1)) return KFCU_EINVAL;
    if (batch <= 0) return 0;
    RealPassParams<AT> S;
    S.in = (const CT*)d_in;
    S.out = (CT*)d_out;
    S.batch = batch;
    S.in_dist = in_dist;
    S.out_dist = out_dist;
    S.nc = plan->nfft;
    S.post = post;
    S.stw = (const CT*)plan->d_stw;
    const long long work = batch * (plan->nfft / 2 + 1);
    long long grid = (work + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (grid > cap) grid = cap;
    kf_realpass_kernel<AT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(S);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_transpose(const void* d_in, void* d_out, long long rows, long long cols, void* stream)
{
    if (!d_in || !d_out || rows < 0 || cols < 0) return KFCU_EINVAL;
    if (rows == 0 || cols == 0) return 0;
    const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    long long grid = (long long)device_info().sms * 8;
    if (grid > tiles) grid = tiles;
    kf_transpose_kernel<CT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const CT*)d_in, (CT*)d_out, rows, cols);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_transpose_peers(const void* d_in, long long in_pitch, void* const* peers, int npeers, long long rows,
                                    long long cols_per_peer, long long out_pitch, long long out_off, void* stream)
{
    if (!d_in || !peers || npeers < 1 || npeers > 16 || rows < 0 || cols_per_peer < 0) return KFCU_EINVAL;
    if (rows == 0 || cols_per_peer == 0) return 0;
    KfPeerPtrs pp;
    for (int s = 0; s < 16; ++s) pp.p[s] = s < npeers ? peers[s] : nullptr;
    const long long tiles = ((rows + 31) / 32) * ((cols_per_peer + 31) / 32) * npeers;
    long long grid = (long long)device_info().sms * 8;
    if (grid > tiles) grid = tiles;
    kf_transpose_peers_kernel<CT><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const CT*)d_in, in_pitch, pp, npeers, rows, cols_per_peer,
                                                                                  out_pitch, out_off);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

extern "C" int kfcu_has_fused(int nfft, int mode)
{
    if (mode < 0 || mode > 3) return 0;
    for (const FusedEntry& e : kFusedTable)
        if (e.N == nfft && e.fn[mode]) return 1;
    return 0;
}

extern "C" int kfcu_generic_max_nfft(void)
{
    const DeviceInfo di = device_info();
    int smem = di.max_smem_optin > 0 ? di.max_smem_optin : 227 * 1024;
    return (int)(smem / (2 * sizeof(CT)));
}

extern "C" long long kfcu_launch_count(void) { return g_launches.load(); }
extern "C" void kfcu_force_generic(int on) { g_force_generic.store(on); }
extern "C" void kfcu_set_grid_limit(int max_ctas) { g_grid_limit = max_ctas > 0 ? max_ctas : 0; }
extern "C" void kfcu_set_sm_reserve(int sms) { g_sm_reserve = sms > 0 ? sms : 0; }
