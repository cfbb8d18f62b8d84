/*
 * cpu_driver.c -- batch driver for timing / running a kissfft-API library on the host CPU.
 *
 * TEST INFRASTRUCTURE ONLY (lives under oracle/): used by tests/ to run the compiled reference over a
 * batch, and by bench.py's cpu_baseline / --impl reference legs to time it.  Never used by the product.
 *
 * It is datatype-agnostic: the library is dlopen()ed by path (RTLD_LOCAL, so several datatype builds that
 * export the same symbol names can coexist in one process) and data is addressed in bytes.  The loop over
 * the batch mirrors the reference's own bench loop (test/benchkiss.c:83-115) with one difference: each
 * iteration works on its own rows instead of re-transforming one zeroed buffer, and rows may be spread over
 * OpenMP threads (cfgs are read-only for kiss_fft -- reference README.md:217 -- and we allocate one cfg per
 * thread because kiss_fftr/kiss_fftnd cfgs embed scratch buffers).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <omp.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

typedef void *(*alloc1_fn)(int, int, void *, size_t *);
typedef void *(*allocnd_fn)(const int *, int, int, void *, size_t *);
typedef void (*xform_fn)(void *, const void *, void *);

enum { K_FFT = 0, K_FFTR = 1, K_FFTRI = 2, K_FFTND = 3, K_FFTNDR = 4, K_FFTNDRI = 5 };

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int cpudrv_max_threads(void) { return omp_get_max_threads(); }

/*
 * Runs `howmany` transforms of the given kind; row b reads in + b*in_dist_bytes and writes
 * out + b*out_dist_bytes.  nthreads <= 1: plain serial loop (so a reference built with KISSFFT_OPENMP can
 * use its own internal parallelism); nthreads > 1: rows are divided over that many OpenMP threads.
 * Returns the best wall-clock time of `reps` repetitions in seconds, or a negative value on error.
 */
double cpudrv_run(const char *libpath, int kind, const int *dims, int ndims, int inverse, const void *in, void *out,
                  size_t howmany, size_t in_dist_bytes, size_t out_dist_bytes, int nthreads, int reps)
{
    void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "cpudrv: dlopen(%s): %s\n", libpath, dlerror()); return -1.0; }
    static const char *alloc_names[] = { "kiss_fft_alloc", "kiss_fftr_alloc", "kiss_fftr_alloc",
                                         "kiss_fftnd_alloc", "kiss_fftndr_alloc", "kiss_fftndr_alloc" };
    static const char *xform_names[] = { "kiss_fft", "kiss_fftr", "kiss_fftri", "kiss_fftnd", "kiss_fftndr", "kiss_fftndri" };
    if (kind < 0 || kind > K_FFTNDRI) return -2.0;
    void *alloc_sym = dlsym(h, alloc_names[kind]);
    xform_fn xform = (xform_fn)dlsym(h, xform_names[kind]);
    if (!alloc_sym || !xform) { fprintf(stderr, "cpudrv: missing symbol in %s\n", libpath); return -3.0; }
    if (nthreads < 1) nthreads = 1;

    void **cfgs = (void **)calloc((size_t)nthreads, sizeof(void *));
    int bad = 0;
    for (int t = 0; t < nthreads; ++t) {
        cfgs[t] = (kind <= K_FFTRI) ? ((alloc1_fn)alloc_sym)(dims[0], inverse, NULL, NULL)
                                    : ((allocnd_fn)alloc_sym)(dims, ndims, inverse, NULL, NULL);
        if (!cfgs[t]) bad = 1;
    }
    double best = -4.0;
    if (!bad) {
        for (int r = 0; r < reps; ++r) {
            double t0 = now_s();
            if (nthreads == 1) {
                for (size_t b = 0; b < howmany; ++b)
                    xform(cfgs[0], (const char *)in + b * in_dist_bytes, (char *)out + b * out_dist_bytes);
            } else {
#pragma omp parallel num_threads(nthreads)
                {
                    void *cfg = cfgs[omp_get_thread_num()];
#pragma omp for schedule(static)
                    for (long long b = 0; b < (long long)howmany; ++b)
                        xform(cfg, (const char *)in + (size_t)b * in_dist_bytes, (char *)out + (size_t)b * out_dist_bytes);
                }
            }
            double dt = now_s() - t0;
            if (best < 0 || dt < best) best = dt;
        }
    }
    for (int t = 0; t < nthreads; ++t) free(cfgs[t]); /* kiss_fft_free == free (kiss_fft.h:138) */
    free(cfgs);
    /* the handle is intentionally kept open: dlclose of an OpenMP-using library is unsafe */
    return best;
}
