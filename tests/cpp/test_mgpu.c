/*
 * test_mgpu.c -- the multi-GPU kiss_fftnd through the C-ABI alone (no Python, no torch): one process per GPU.
 *
 *   test_mgpu G d0 d1 d2 [flags] [iters]          flags: 1 = peer-store exchange, 2 = reference axis order (exact mode)
 *                                                 d1 = 1: the 2-D transform d0 x d2
 *
 * Builds for every datatype (-DFIXED_POINT=16|32, -Dkiss_fft_scalar=...): the fixed-point builds must reproduce the
 * single-GPU kiss_fftnd bit for bit in the reference-order mode.
 *
 * The parent forks G children before any CUDA call.  Child 0 asks the library for the rendezvous id and hands it to the
 * others through pipes; every child selects GPU `rank`, transforms its slab with kiss_fftnd_mgpu_exec and compares the
 * transposed-out result with the single-GPU kiss_fftnd_dev of the whole array (computed on its own GPU, which is why the
 * test sizes are small).  Exit status 0 = every rank within 2e-6*log2(N) relative RMS.  With iters > 0 it also prints the
 * fenced per-call time (cudaEvent around one exec, MAX over... each rank prints its own mean).
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include "kiss_fft_cuda.h"

#define CK(x)                                                                                   \
    do {                                                                                        \
        int e_ = (int)(x);                                                                      \
        if (e_ != 0) { fprintf(stderr, "rank %d: %s failed (%d) %s\n", rank, #x, e_, kiss_fftnd_mgpu_last_error()); return 10; } \
    } while (0)

static kiss_fft_scalar rnd(unsigned long long *s)
{
    *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
    const double u = (double)(*s & 0xffffff) / 0x1000000 * 2.0 - 1.0;
#ifdef FIXED_POINT      /* uniform in +-SAMP_MAX/2: the range where the reference is compiler-independent */
    return (kiss_fft_scalar)floor(u * (FIXED_POINT == 16 ? 16383.0 : 1073741823.0));
#else
    return (kiss_fft_scalar)u;
#endif
}

static int run_rank(int rank, int G, const int *dims, unsigned flags, int iters, const void *id)
{
    const size_t d0 = dims[0], d1 = dims[1], d2 = dims[2], n = d0 * d1 * d2, P = d0 / G, C = d2 / G;
    /* d1 == 1 on the command line = the 2-D transform d0 x d2 (kiss_fftnd_mgpu_alloc with ndims = 2) */
    const int dims2[2] = {dims[0], dims[2]};
    const int nd_n = dims[1] == 1 ? 2 : 3;
    const int *nd_dims = nd_n == 2 ? dims2 : dims;
    CK(cudaSetDevice(rank));
    if (getenv("MGPU_TEST_TIMING_ONLY")) {      /* tuning aid: no host arrays, no verification, just the fenced timing */
        kiss_fftnd_mgpu_cfg c2 = kiss_fftnd_mgpu_alloc(nd_dims, nd_n, 0, rank, G, id, flags);
        if (!c2) { fprintf(stderr, "rank %d: alloc failed: %s\n", rank, kiss_fftnd_mgpu_last_error()); return 11; }
        kiss_fft_cpx *a, *b;
        CK(cudaMalloc((void **)&a, sizeof(kiss_fft_cpx) * P * d1 * d2));
        CK(cudaMalloc((void **)&b, sizeof(kiss_fft_cpx) * C * d1 * d0));
        CK(cudaMemset(a, 0, sizeof(kiss_fft_cpx) * P * d1 * d2));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        /* MGPU_TEST_SWEEP="chunks:pchunks:b_ctas:prio:ac_reserve:trace:link_sms:tail16,..." : every entry is applied with kiss_fftnd_mgpu_tune and timed */
        const char *sweep = getenv("MGPU_TEST_SWEEP");
        char buf[1024];
        snprintf(buf, sizeof(buf), "%s", sweep && *sweep ? sweep : "-1:-1:-1:-1");
        for (char *save = NULL, *tok = strtok_r(buf, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
            int k[9] = {-1, -1, -1, -1, -1, 0, -1, -1, -1};
            sscanf(tok, "%d:%d:%d:%d:%d:%d:%d:%d", &k[0], &k[1], &k[2], &k[3], &k[4], &k[5], &k[6], &k[8]);
            CK(kiss_fftnd_mgpu_tune(c2, k, 9));
            for (int i = 0; i < 3; ++i) CK(kiss_fftnd_mgpu_exec(c2, a, b, NULL));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, NULL));
            for (int i = 0; i < iters; ++i) CK(kiss_fftnd_mgpu_exec(c2, a, b, NULL));
            CK(cudaEventRecord(e1, NULL));
            CK(cudaEventSynchronize(e1));
            float t = 0;
            CK(cudaEventElapsedTime(&t, e0, e1));
            if (rank == 0 || getenv("MGPU_TEST_ALL_RANKS"))
                printf("{\"rank\": %d, \"ranks\": %d, \"p2p\": %d, \"chunks\": %d, \"pchunks\": %d, \"b_ctas\": %d, \"prio\": %d, \"ac_reserve\": %d, \"link_sms\": %d, \"rest_sms\": %d, \"last_chunk_cols\": %d, \"ms\": %.4f}\n", rank, G,
                       kiss_fftnd_mgpu_uses_p2p(c2), kiss_fftnd_mgpu_knob(c2, 0), kiss_fftnd_mgpu_knob(c2, 1), kiss_fftnd_mgpu_knob(c2, 2),
                       kiss_fftnd_mgpu_knob(c2, 3), kiss_fftnd_mgpu_knob(c2, 4), kiss_fftnd_mgpu_knob(c2, 6), kiss_fftnd_mgpu_knob(c2, 7), kiss_fftnd_mgpu_knob(c2, 8), t / (iters > 0 ? iters : 1));
            fflush(stdout);
            if (k[5] > 0 && rank == 0) {       /* sixth field: print the launch timeline of the last exec */
                static char tl[8192];
                CK(kiss_fftnd_mgpu_trace(c2, tl, sizeof(tl)));
                printf("%s", tl);
                fflush(stdout);
            }
        }
        kiss_fftnd_mgpu_free(c2);
        return 0;
    }
    kiss_fft_cpx *h = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * n);
    unsigned long long s = 88172645463325252ULL;
    for (size_t i = 0; i < n; ++i) { h[i].r = rnd(&s); h[i].i = rnd(&s); }
    kiss_fft_cpx *d_full, *d_ref, *d_in, *d_out;
    CK(cudaMalloc((void **)&d_full, sizeof(kiss_fft_cpx) * n));
    CK(cudaMalloc((void **)&d_ref, sizeof(kiss_fft_cpx) * n));
    CK(cudaMemcpy(d_full, h, sizeof(kiss_fft_cpx) * n, cudaMemcpyHostToDevice));
    kiss_fftnd_cfg nd = kiss_fftnd_alloc(nd_dims, nd_n, 0, NULL, NULL);
    CK(kiss_fftnd_dev(nd, d_full, d_ref, NULL, NULL));
    CK(cudaDeviceSynchronize());
    kiss_fft_cpx *ref = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * n);
    CK(cudaMemcpy(ref, d_ref, sizeof(kiss_fft_cpx) * n, cudaMemcpyDeviceToHost));

    kiss_fftnd_mgpu_cfg cfg = kiss_fftnd_mgpu_alloc(nd_dims, nd_n, 0, rank, G, id, flags);
    if (!cfg) { fprintf(stderr, "rank %d: kiss_fftnd_mgpu_alloc failed: %s\n", rank, kiss_fftnd_mgpu_last_error()); return 11; }
    const size_t nin = kiss_fftnd_mgpu_local_in_elems(cfg), nout = kiss_fftnd_mgpu_local_out_elems(cfg);
    if (nin != P * d1 * d2 || nout != C * d1 * d0) { fprintf(stderr, "rank %d: slab sizes\n", rank); return 12; }
    CK(cudaMalloc((void **)&d_in, sizeof(kiss_fft_cpx) * nin));
    CK(cudaMalloc((void **)&d_out, sizeof(kiss_fft_cpx) * nout));
    const int reford = (flags & KISS_FFT_MGPU_REFERENCE_ORDER) != 0;
    kiss_fft_cpx *hslab = NULL;
    if (reford) {                            /* this rank's slab along the last axis: [d0][d1][C] */
        hslab = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * nin);
        for (size_t a = 0; a < d0 * d1; ++a) memcpy(hslab + a * C, h + a * d2 + (size_t)rank * C, sizeof(kiss_fft_cpx) * C);
    }
    double worst = 0;
    long long mismatches = 0;
    for (int rep = 0; rep < 2; ++rep) {      /* twice: the second call exercises the buffer-reuse handshake */
        CK(cudaMemcpy(d_in, reford ? hslab : h + (size_t)rank * nin, sizeof(kiss_fft_cpx) * nin, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_out, 0, sizeof(kiss_fft_cpx) * nout));
        CK(kiss_fftnd_mgpu_exec(cfg, d_in, d_out, NULL));
        CK(cudaDeviceSynchronize());
        kiss_fft_cpx *got = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * nout);
        CK(cudaMemcpy(got, d_out, sizeof(kiss_fft_cpx) * nout, cudaMemcpyDeviceToHost));
        double num = 0, den = 0;
        if (reford) {                        /* natural-order rows rank*P .. of kiss_fftnd's output, compared bit for bit too */
            const kiss_fft_cpx *w = ref + (size_t)rank * nout;
            for (size_t i = 0; i < nout; ++i) {
                num += ((double)got[i].r - w[i].r) * ((double)got[i].r - w[i].r) + ((double)got[i].i - w[i].i) * ((double)got[i].i - w[i].i);
                den += (double)w[i].r * w[i].r + (double)w[i].i * w[i].i;
                mismatches += (got[i].r != w[i].r) || (got[i].i != w[i].i);
            }
        } else
        for (size_t c = 0; c < C; ++c)
            for (size_t k1 = 0; k1 < d1; ++k1)
                for (size_t k0 = 0; k0 < d0; ++k0) {
                    const kiss_fft_cpx g = got[(c * d1 + k1) * d0 + k0], w = ref[(k0 * d1 + k1) * d2 + rank * C + c];
                    num += ((double)g.r - w.r) * ((double)g.r - w.r) + ((double)g.i - w.i) * ((double)g.i - w.i);
                    den += (double)w.r * w.r + (double)w.i * w.i;
                }
        const double err = sqrt(num / den);
        if (err > worst) worst = err;
        free(got);
    }
#ifdef FIXED_POINT
    const double tol = reford ? 0.0 : 1.0;        /* exact mode: bit-identical; the fast order is not defined for fixed point */
#else
    const double tol = (sizeof(kiss_fft_scalar) == 8 ? 2e-14 : 2e-6) * log2((double)n);
#endif
    float ms = 0;
    if (iters > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) CK(kiss_fftnd_mgpu_exec(cfg, d_in, d_out, NULL));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, NULL));
        for (int i = 0; i < iters; ++i) CK(kiss_fftnd_mgpu_exec(cfg, d_in, d_out, NULL));
        CK(cudaEventRecord(e1, NULL));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= (float)iters;
    }
    printf("{\"rank\": %d, \"ranks\": %d, \"dims\": [%zu, %zu, %zu], \"p2p\": %d, \"chunks\": %d, \"reference_order\": %d, \"mismatches\": %lld, "
           "\"rel_rms\": %.3g, \"tol\": %.3g, \"ms\": %.4f}\n",
           rank, G, d0, d1, d2, kiss_fftnd_mgpu_uses_p2p(cfg), kiss_fftnd_mgpu_chunks(cfg), reford, mismatches, worst, tol, ms);
    fflush(stdout);
    kiss_fftnd_mgpu_free(cfg);
    free(nd);
#ifdef FIXED_POINT
    if (reford && mismatches != 0) return 21;
#endif
    return worst <= tol ? 0 : 20;
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s G d0 d1 d2 [flags] [iters]\n", argv[0]); return 2; }
    const int G = atoi(argv[1]);
    const int dims[3] = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
    const unsigned flags = argc > 5 ? (unsigned)atoi(argv[5]) : 0;      /* KISS_FFT_MGPU_P2P | KISS_FFT_MGPU_REFERENCE_ORDER */
    const int iters = argc > 6 ? atoi(argv[6]) : 0;
    if (G < 1 || G > 16) return 2;
    if (G == 1) return run_rank(0, 1, dims, flags, iters, NULL);
    int pipes[16][2];
    for (int r = 1; r < G; ++r)
        if (pipe(pipes[r]) != 0) return 3;
    pid_t pid[16];
    for (int r = 0; r < G; ++r) {
        pid[r] = fork();
        if (pid[r] == 0) {
            char id[KISS_FFT_MGPU_ID_BYTES];
            if (r == 0) {
                if (kiss_fftnd_mgpu_get_id(id) != 0) { fprintf(stderr, "get_id: %s\n", kiss_fftnd_mgpu_last_error()); _exit(4); }
                for (int q = 1; q < G; ++q)
                    if (write(pipes[q][1], id, sizeof(id)) != (ssize_t)sizeof(id)) _exit(5);
            } else if (read(pipes[r][0], id, sizeof(id)) != (ssize_t)sizeof(id)) {
                _exit(6);
            }
            _exit(run_rank(r, G, dims, flags, iters, id));
        }
    }
    int bad = 0;
    for (int r = 0; r < G; ++r) {
        int st = 0;
        waitpid(pid[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "rank %d exited with %d\n", r, WIFEXITED(st) ? WEXITSTATUS(st) : -1); bad = 1; }
    }
    return bad;
}
