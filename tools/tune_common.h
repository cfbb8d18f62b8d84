// tune_common.h -- harness shared by the generated tuner translation units (development aid, see tune_gen.py)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../kissfft_b200/csrc/kf_twtab.h"
#include "../kissfft_b200/csrc/kf_tmap.h"

typedef kf::Arith<kiss_fft_scalar> AT;
typedef AT::C CT;

struct TuneEntry {
    long long (*rows_ok)(const kf::KParams<AT>&);
    std::string label;
    int threads, tpc;
    size_t smem;
    const void* kernel;
    void (*launch)(const kf::KParams<AT>&, unsigned grid, size_t smem);
    void (*prepare)(kf::KParams<AT>&, const CT* h_tw, CT** d_gtw);
};

template <class PT, int MODE>
static void launch_variant(const kf::KParams<AT>& P, unsigned grid, size_t smem)
{
    kf::kf_fused_kernel<AT, PT, MODE><<<grid, PT::D.threads(), smem>>>(P);
}

template <class PT, int MODE>
static void prepare_variant(kf::KParams<AT>& P, const CT* h_tw, CT** d_gtw)
{
    if constexpr (kf::FusedLayout<AT, PT, MODE>::kColRing) {
        if (!kf::col_ring_ok<AT, PT>(P) || kf::col_ring_encode<AT, PT, MODE>(P) != 0) { fprintf(stderr, "tensor map not encodable\n"); exit(3); }
    }
    std::vector<CT> tab = kf::build_gtw<AT, PT>(h_tw);
    if (*d_gtw) cudaFree(*d_gtw);
    cudaMalloc(d_gtw, tab.size() * sizeof(CT));
    cudaMemcpy(*d_gtw, tab.data(), tab.size() * sizeof(CT), cudaMemcpyHostToDevice);
    P.gtw = *d_gtw;
    kf::fill_g0tw<AT, PT>(P, h_tw);
}

template <class PT, int MODE>
static TuneEntry make_entry(const char* label)
{
    constexpr kf::PlanDesc D = PT::D;
    TuneEntry e;
    e.label = label;
    e.threads = D.threads();
    e.tpc = D.tpc;
    e.smem = kf::FusedLayout<AT, PT, MODE>::kTotal;
    e.kernel = (const void*)kf::kf_fused_kernel<AT, PT, MODE>;
    e.launch = launch_variant<PT, MODE>;
    e.prepare = prepare_variant<PT, MODE>;
    e.rows_ok = kf::fused_rows<AT, PT, MODE>;
    return e;
}

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

static kiss_fft_scalar from_double(double x)
{
#ifdef FIXED_POINT
    const double smax = (sizeof(kiss_fft_scalar) == 2) ? 32767.0 : 2147483647.0;
    return (kiss_fft_scalar)floor(.5 + smax * x);
#else
    return (kiss_fft_scalar)x;
#endif
}

static int tune_main(int argc, char** argv, std::vector<TuneEntry>& vars, int nfft, int mode)
{
    const long long batch = argc > 1 ? atoll(argv[1]) : 65536;
    const int iters = argc > 2 ? atoi(argv[2]) : 10;
    const char* filter = argc > 3 ? argv[3] : nullptr;
    CT* d_gtw = nullptr;
    const bool real_in = (mode == kf::kR2C), real_out = (mode == kf::kC2R);
    // element counts per row in complex units (real rows are packed complex of nfft)
    const long long in_row = real_out ? nfft + 1 : nfft, out_row = real_in ? nfft + 1 : nfft;
    std::vector<CT> h_tw(nfft), h_stw(nfft / 2 + 1);
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    const int inverse = real_out ? 1 : 0;
    for (int i = 0; i < nfft; ++i) {
        double ph = -2 * pi * i / nfft;
        if (inverse) ph = -ph;
        h_tw[i].r = from_double(cos(ph));
        h_tw[i].i = from_double(sin(ph));
    }
    for (int i = 0; i < nfft / 2; ++i) {
        double ph = -3.14159265358979323846264338327 * ((double)(i + 1) / nfft + .5);
        if (inverse) ph = -ph;
        h_stw[i].r = from_double(cos(ph));
        h_stw[i].i = from_double(sin(ph));
    }
    CT *d_in, *d_out, *d_ref, *d_tw, *d_stw;
    CK(cudaMalloc(&d_in, sizeof(CT) * in_row * batch));
    CK(cudaMalloc(&d_out, sizeof(CT) * out_row * batch));
    CK(cudaMalloc(&d_ref, sizeof(CT) * out_row * batch));
    CK(cudaMalloc(&d_tw, sizeof(CT) * nfft));
    CK(cudaMalloc(&d_stw, sizeof(CT) * (nfft / 2 + 1)));
    CK(cudaMemcpy(d_tw, h_tw.data(), sizeof(CT) * nfft, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_stw, h_stw.data(), sizeof(CT) * (nfft / 2 + 1), cudaMemcpyHostToDevice));
    {
        std::vector<CT> h((size_t)in_row * batch);
        unsigned long long s = 88172645463325252ULL;
        for (auto& c : h) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            c.r = from_double(((double)(s & 0xffffff) / 0x1000000 - 0.5));
            c.i = from_double(((double)((s >> 24) & 0xffffff) / 0x1000000 - 0.5));
        }
        CK(cudaMemcpy(d_in, h.data(), sizeof(CT) * h.size(), cudaMemcpyHostToDevice));
    }
    kf::KParams<AT> P;
    P.ncols = 0;
    P.in_pdist = P.out_pdist = 0;
    P.npeers = 0;
    P.cols_per_peer = 0;
    P.peer_col_dist = 0;
    P.max_ctas = 0;
    P.in = d_in; P.out = d_out; P.howmany = batch;
    P.in_dist = in_row; P.out_dist = out_row; P.in_stride = 1;
    if (mode == kf::kC2CCol || mode == kf::kC2CColCol) {
        // TUNE_NCOLS=c: planes of c columns (row stride c elements, like an inner axis of an N-D array); default: one plane
        const char* nc = getenv("TUNE_NCOLS");
        const long long ncols = nc ? atoll(nc) : 0;
        P.in_dist = 1;
        P.in_stride = ncols > 0 ? ncols : batch;
        P.out_dist = (mode == kf::kC2CCol) ? nfft : 1;
        if (ncols > 0) { P.ncols = ncols; P.in_pdist = (long long)nfft * ncols; P.out_pdist = (long long)nfft * ncols; }
    }
    P.tw = d_tw; P.stw = d_stw;
    CT z{};
    P.pc.epi3 = AT::load(nfft % 3 == 0 ? h_tw[nfft / 3] : z);
    P.pc.ya = AT::load(nfft % 5 == 0 ? h_tw[nfft / 5] : z);
    P.pc.yb = AT::load(nfft % 5 == 0 ? h_tw[2 * (nfft / 5)] : z);
    P.inverse = inverse;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t esz = sizeof(CT);
    const double bytes = (double)batch * esz * (in_row + out_row) - (real_in || real_out ? (double)batch * esz : 0.0);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<CT> h_ref((size_t)out_row * 64), h_out((size_t)out_row * 64);
    bool have_ref = false;
    for (auto& v : vars) {
        if (filter && v.label.find(filter) == std::string::npos) continue;
        v.prepare(P, h_tw.data(), &d_gtw);
        if (v.rows_ok(P) != P.howmany) { printf("{\"variant\": \"%s\", \"error\": \"alignment rules\"}\n", v.label.c_str()); continue; }
        if (cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem) != cudaSuccess) {
            cudaGetLastError();
            printf("{\"variant\": \"%s\", \"error\": \"smem %zu too large\"}\n", v.label.c_str(), v.smem);
            continue;
        }
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, v.kernel, v.threads, v.smem));
        if (nb < 1) { printf("{\"variant\": \"%s\", \"error\": \"does not fit\"}\n", v.label.c_str()); continue; }
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, v.kernel));
        const long long ntiles = (batch + v.tpc - 1) / v.tpc;
        const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)sms * nb);
        P.out = d_out;
        CK(cudaMemset(d_out, 0, sizeof(CT) * out_row * batch));
        for (int i = 0; i < 3; ++i) v.launch(P, grid, v.smem);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<float> ms(iters);
        for (int i = 0; i < iters; ++i) {
            CK(cudaEventRecord(e0));
            v.launch(P, grid, v.smem);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms[i], e0, e1));
        }
        std::sort(ms.begin(), ms.end());
        // consistency with the first working variant (last 64 rows)
        CK(cudaMemcpy(h_out.data(), d_out + (size_t)out_row * (batch - 64), sizeof(CT) * h_out.size(), cudaMemcpyDeviceToHost));
        double maxdiff = 0;
        if (!have_ref) { h_ref = h_out; have_ref = true; }
        else for (size_t i = 0; i < h_out.size(); ++i) {
            maxdiff = std::max(maxdiff, fabs((double)h_out[i].r - (double)h_ref[i].r));
            maxdiff = std::max(maxdiff, fabs((double)h_out[i].i - (double)h_ref[i].i));
        }
        const double med = ms[iters / 2], best = ms[0];
        printf("{\"variant\": \"%s\", \"ms_med\": %.4f, \"ms_best\": %.4f, \"gbs_med\": %.1f, \"gbs_best\": %.1f, \"regs\": %d, "
               "\"ctas_per_sm\": %d, \"threads\": %d, \"smem\": %zu, \"maxdiff\": %.3g}\n",
               v.label.c_str(), med, best, bytes / med * 1e-6, bytes / best * 1e-6, fa.numRegs, nb, v.threads, v.smem, maxdiff);
        fflush(stdout);
    }
    return 0;
}
