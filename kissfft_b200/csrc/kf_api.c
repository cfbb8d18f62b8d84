/*
 * kf_api.c -- host side of kissfft-b200, plain C.
 *
 * Implements the reference's public API (include/kiss_fft.h, kiss_fftr.h, kiss_fftnd.h, kiss_fftndr.h) and
 * the batched device-pointer extension (include/kiss_fft_cuda.h) on top of the CUDA launchers of
 * kf_launch.cu.  What lives here:
 *   - the planner: radix schedule and twiddle tables, generated on the host in double precision with the
 *     reference's exact expressions so that the Q15/Q31 tables are bit-identical
 *     (kf_factor kiss_fft.c:306-328, twiddles kiss_fft.c:361-367, split twiddles kiss_fftr.c:53-59);
 *   - the cfg objects: single free()-able POD blocks honouring the mem/lenmem placement protocol
 *     (kiss_fft.h:94-115) -- they contain no device handles, so they may be copied or freed at will;
 *   - a per-(device, nfft, direction) cache of device-side tables, released by kiss_fft_cleanup();
 *   - pointer classification: host pointers are staged through the GPU, device pointers run in place;
 *   - the axis-pass orchestration of kiss_fftnd / kiss_fftndr (kiss_fftnd.c:156-188, kiss_fftndr.c:86-132).
 * There is no CPU transform code in this file or anywhere in the library.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/kfc.h"
#include "../../include/kiss_fft_cuda.h"
#include "kf_internal.h"

#ifdef FIXED_POINT
# if (FIXED_POINT == 32)
#  define KF_SAMP_MAX 2147483647
# else
#  define KF_SAMP_MAX 32767
# endif
#endif

#define KF_MAGIC_1D 0x4b463144u
#define KF_MAGIC_R 0x4b465245u
#define KF_MAGIC_ND 0x4b464e44u
#define KF_MAGIC_NDR 0x4b464e52u

struct kiss_fft_state {
    uint32_t magic;
    int nfft;
    int inverse;
    int nstages;
    int factors[2 * KFCU_MAXSTAGES]; /* p0,m0,p1,m1,... like the reference's factors[] */
    kiss_fft_cpx twiddles[1];        /* nfft entries */
};

struct kiss_fftr_state {
    uint32_t magic;
    int nfft; /* real length */
    kiss_fft_cfg substate;
    kiss_fft_cpx *super_twiddles; /* nfft/4 entries */
};

struct kiss_fftnd_state {
    uint32_t magic;
    int ndims;
    int inverse;
    long long dimprod;
    int *dims;
    kiss_fft_cfg *states;
};

struct kiss_fftndr_state {
    uint32_t magic;
    int dimReal;
    long long dimOther;
    int ndims;
    int inverse;
    kiss_fftr_cfg cfg_r;
    kiss_fftnd_cfg cfg_nd; /* over dims[0..ndims-2]; NULL when ndims == 1 */
};

/* ---- error reporting (reference convention: "[ERROR] file:line msg" on stderr, kiss_fft_log.h:20-32) ---- */
static __thread char tls_err[256];

static void kf_error(const char *file, int line, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_err, sizeof(tls_err), fmt, ap);
    va_end(ap);
#ifndef NDEBUG
    fprintf(stderr, "[ERROR] %s:%d %s\n", file, line, tls_err);
#endif
}
#define KF_ERROR(...) kf_error(__FILE__, __LINE__, __VA_ARGS__)

static int kf_cuda_fail(const char *file, int line, const char *what, int code)
{
    if (code > 0)
        kf_error(file, line, "%s: CUDA error %d (%s)", what, code, cudaGetErrorString((cudaError_t)code));
    else
        kf_error(file, line, "%s: error %d", what, code);
    return code;
}
#define KF_CHECK(expr)                                                    \
    do {                                                                  \
        int kf_rc_ = (int)(expr);                                         \
        if (kf_rc_ != 0) return kf_cuda_fail(__FILE__, __LINE__, #expr, kf_rc_); \
    } while (0)

const char *kiss_fft_cuda_last_error(void) { return tls_err; }
long long kiss_fft_cuda_launch_count(void) { return kfcu_launch_count(); }
int kiss_fft_cuda_scalar_bytes(void) { return (int)sizeof(kiss_fft_scalar); }
int kiss_fft_cuda_is_fixed_point(void)
{
#ifdef FIXED_POINT
    return 1;
#else
    return 0;
#endif
}

/* ---- planner -------------------------------------------------------------------------------------------- */

/* trig -> scalar, the reference's KISS_FFT_COS/SIN (_kiss_fft_guts.h:127-139) */
static kiss_fft_scalar kf_scalar_from(double x)
{
#ifdef FIXED_POINT
    return (kiss_fft_scalar)floor(.5 + KF_SAMP_MAX * x);
#else
    return (kiss_fft_scalar)x;
#endif
}

/* 4s first, then 2s, then 3, 5, 7, ...; once p exceeds floor(sqrt(n_original)) the remainder is prime
 * (kiss_fft.c:306-328).  Returns the number of stages. */
static int kf_plan_radices(int n, int *facbuf)
{
    int p = 4, ns = 0;
    const double root = floor(sqrt((double)n));
    do {
        while (n % p) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if (p > root) p = n;
        }
        n /= p;
        if (ns >= KFCU_MAXSTAGES) return -1;
        facbuf[2 * ns] = p;
        facbuf[2 * ns + 1] = n;
        ++ns;
    } while (n > 1);
    return ns;
}

kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem)
{
    kiss_fft_cfg st = NULL;
    if (nfft <= 0) return NULL;
    if ((size_t)nfft >= (SIZE_MAX - 2 * sizeof(struct kiss_fft_state)) / sizeof(kiss_fft_cpx)) return NULL;
    const size_t memneeded = sizeof(struct kiss_fft_state) + sizeof(kiss_fft_cpx) * (size_t)(nfft - 1);

    if (lenmem == NULL) {
        st = (kiss_fft_cfg)KISS_FFT_MALLOC(memneeded);
    } else {
        if (mem != NULL && *lenmem >= memneeded) st = (kiss_fft_cfg)mem;
        *lenmem = memneeded;
    }
    if (!st) return NULL;

    st->magic = KF_MAGIC_1D;
    st->nfft = nfft;
    st->inverse = inverse_fft ? 1 : 0;
    memset(st->factors, 0, sizeof(st->factors));
    st->nstages = kf_plan_radices(nfft, st->factors);
    {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        for (int i = 0; i < nfft; ++i) {
            double phase = -2 * pi * i / nfft;
            if (st->inverse) phase *= -1;
            st->twiddles[i].r = kf_scalar_from(cos(phase));
            st->twiddles[i].i = kf_scalar_from(sin(phase));
        }
    }
    return st;
}

int kiss_fft_next_fast_size(int n)
{
    for (;; ++n) {
        int m = n;
        while ((m % 2) == 0) m /= 2;
        while ((m % 3) == 0) m /= 3;
        while ((m % 5) == 0) m /= 5;
        if (m <= 1) return n;
    }
}

kiss_fftr_cfg kiss_fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem)
{
    kiss_fftr_cfg st = NULL;
    size_t subsize = 0;
    if (nfft & 1) {
        KF_ERROR("Real FFT optimization must be even.");
        return NULL;
    }
    if (nfft <= 0) return NULL;
    const int nc = nfft >> 1;
    kiss_fft_alloc(nc, inverse_fft, NULL, &subsize);
    subsize = (subsize + 15u) & ~(size_t)15u;
    const size_t hdr = (sizeof(struct kiss_fftr_state) + 15u) & ~(size_t)15u;
    const size_t memneeded = hdr + subsize + sizeof(kiss_fft_cpx) * (size_t)(nc / 2 + 1);

    if (lenmem == NULL) {
        st = (kiss_fftr_cfg)KISS_FFT_MALLOC(memneeded);
    } else {
        if (mem != NULL && *lenmem >= memneeded) st = (kiss_fftr_cfg)mem;
        *lenmem = memneeded;
    }
    if (!st) return NULL;

    st->magic = KF_MAGIC_R;
    st->nfft = nfft;
    st->substate = (kiss_fft_cfg)((char *)st + hdr);
    st->super_twiddles = (kiss_fft_cpx *)((char *)st->substate + subsize);
    kiss_fft_alloc(nc, inverse_fft, st->substate, &subsize);
    for (int i = 0; i < nc / 2; ++i) {
        double phase = -3.14159265358979323846264338327 * ((double)(i + 1) / nc + .5);
        if (inverse_fft) phase *= -1;
        st->super_twiddles[i].r = kf_scalar_from(cos(phase));
        st->super_twiddles[i].i = kf_scalar_from(sin(phase));
    }
    return st;
}

kiss_fftnd_cfg kiss_fftnd_alloc(const int *dims, int ndims, int inverse_fft, void *mem, size_t *lenmem)
{
    kiss_fftnd_cfg st = NULL;
    if (ndims < 0 || (ndims > 0 && !dims)) return NULL;
    size_t memneeded = (sizeof(struct kiss_fftnd_state) + 15u) & ~(size_t)15u;
    const size_t off_dims = memneeded;
    memneeded += (sizeof(int) * (size_t)ndims + 15u) & ~(size_t)15u;
    const size_t off_states = memneeded;
    memneeded += (sizeof(void *) * (size_t)ndims + 15u) & ~(size_t)15u;
    const size_t off_cfgs = memneeded;
    long long dimprod = 1;
    for (int i = 0; i < ndims; ++i) {
        size_t sublen = 0;
        if (dims[i] <= 0) return NULL;
        kiss_fft_alloc(dims[i], inverse_fft, NULL, &sublen);
        memneeded += (sublen + 15u) & ~(size_t)15u;
        dimprod *= dims[i];
    }
    if (lenmem == NULL) {
        st = (kiss_fftnd_cfg)KISS_FFT_MALLOC(memneeded);
    } else {
        if (mem != NULL && *lenmem >= memneeded) st = (kiss_fftnd_cfg)mem;
        *lenmem = memneeded;
    }
    if (!st) return NULL;
    st->magic = KF_MAGIC_ND;
    st->ndims = ndims;
    st->inverse = inverse_fft ? 1 : 0;
    st->dimprod = dimprod;
    st->dims = (int *)((char *)st + off_dims);
    st->states = (kiss_fft_cfg *)((char *)st + off_states);
    char *ptr = (char *)st + off_cfgs;
    for (int i = 0; i < ndims; ++i) {
        size_t len = 0;
        st->dims[i] = dims[i];
        kiss_fft_alloc(dims[i], inverse_fft, NULL, &len);
        st->states[i] = kiss_fft_alloc(dims[i], inverse_fft, ptr, &len);
        ptr += (len + 15u) & ~(size_t)15u;
    }
    return st;
}

kiss_fftndr_cfg kiss_fftndr_alloc(const int *dims, int ndims, int inverse_fft, void *mem, size_t *lenmem)
{
    kiss_fftndr_cfg st = NULL;
    size_t nr = 0, nd = 0;
    if (ndims < 1 || !dims) return NULL;
    const int dimReal = dims[ndims - 1];
    if (dimReal <= 0 || (dimReal & 1)) {
        KF_ERROR("Real FFT optimization must be even.");
        return NULL;
    }
    long long dimOther = 1;
    for (int i = 0; i < ndims - 1; ++i) {
        if (dims[i] <= 0) return NULL;
        dimOther *= dims[i];
    }
    (void)kiss_fftr_alloc(dimReal, inverse_fft, NULL, &nr);
    if (ndims > 1) (void)kiss_fftnd_alloc(dims, ndims - 1, inverse_fft, NULL, &nd);
    const size_t hdr = (sizeof(struct kiss_fftndr_state) + 15u) & ~(size_t)15u;
    nr = (nr + 15u) & ~(size_t)15u;
    nd = (nd + 15u) & ~(size_t)15u;
    const size_t memneeded = hdr + nr + nd;
    if (lenmem == NULL) {
        st = (kiss_fftndr_cfg)malloc(memneeded);
    } else {
        if (mem != NULL && *lenmem >= memneeded) st = (kiss_fftndr_cfg)mem;
        *lenmem = memneeded;
    }
    if (!st) return NULL;
    memset(st, 0, memneeded);
    st->magic = KF_MAGIC_NDR;
    st->dimReal = dimReal;
    st->dimOther = dimOther;
    st->ndims = ndims;
    st->inverse = inverse_fft ? 1 : 0;
    char *ptr = (char *)st + hdr;
    st->cfg_r = kiss_fftr_alloc(dimReal, inverse_fft, ptr, &nr);
    ptr += nr;
    st->cfg_nd = (ndims > 1) ? kiss_fftnd_alloc(dims, ndims - 1, inverse_fft, ptr, &nd) : NULL;
    return st;
}

/* ---- device plan cache ---------------------------------------------------------------------------------- */
typedef struct kf_devplan {
    int device, nfft, inverse, has_stw;
    kfcu_plan plan;
    kiss_fft_cpx *h_tw; /* private host copy (the cfg may be freed by the caller at any time) */
    struct kf_devplan *next;
} kf_devplan;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static kf_devplan *g_plans = NULL;

/* ---- staging contexts --------------------------------------------------------------------------------------
 * Everything a host-pointer call needs on the device side lives in a staging context: a stream, grow-only device
 * scratch slots and grow-only pinned bounce buffers.  A call takes an idle context of the current device from the pool
 * (creating one when all are busy) and gives it back when its result is in the caller's buffer, so concurrent calls on
 * different threads run concurrently -- the reference's kiss_fft is thread-parallel on a shared cfg
 * (reference README.md:217) -- and the pool lock is held only for the hand-over.  kiss_fft_cleanup() frees the idle
 * contexts; it must not run concurrently with transforms (it also frees the device tables they use).
 *
 * Lock order (never nested otherwise): g_mp_lock -> g_pool_lock -> g_lock. */
typedef struct kf_ctx {
    int device, busy;
    cudaStream_t stream;
    void *d_buf[3];                 /* device scratch: input, output, work */
    size_t d_bytes[3];
    void *h_pin[2];                 /* pinned bounce buffers: input, output */
    size_t h_bytes[2];
    struct kf_ctx *next;
} kf_ctx;

static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;
static kf_ctx *g_pool = NULL;
static pthread_mutex_t g_mp_lock = PTHREAD_MUTEX_INITIALIZER; /* work buffers of the long-row paths (kf_mp_scratch) */
typedef struct { void *ptr; size_t bytes; int device; } kf_buf;
#define KF_MP_SLOTS 3
static kf_buf g_mp_bufs[KF_MP_SLOTS];

static int kf_get_devplan(const struct kiss_fft_state *cfg, const kiss_fft_cpx *stw, const kf_devplan **out)
{
    int dev = 0;
    KF_CHECK(cudaGetDevice(&dev));
    pthread_mutex_lock(&g_lock);
    kf_devplan *e;
    for (e = g_plans; e; e = e->next)
        if (e->device == dev && e->nfft == cfg->nfft && e->inverse == cfg->inverse && (!stw || e->has_stw)) break;
    if (!e) {
        e = (kf_devplan *)calloc(1, sizeof(*e));
        int rc = e ? 0 : KISS_FFT_CUDA_ENOMEM;
        void *d_tw = NULL, *d_stw = NULL;
        const size_t twbytes = sizeof(kiss_fft_cpx) * (size_t)cfg->nfft;
        const size_t stwn = (size_t)(cfg->nfft / 2);
        if (!rc) {
            e->h_tw = (kiss_fft_cpx *)malloc(twbytes);
            if (!e->h_tw) rc = KISS_FFT_CUDA_ENOMEM;
        }
        if (!rc) rc = (int)cudaMalloc(&d_tw, twbytes);
        if (!rc) rc = (int)cudaMemcpy(d_tw, cfg->twiddles, twbytes, cudaMemcpyHostToDevice);
        if (!rc && stw && stwn) {
            rc = (int)cudaMalloc(&d_stw, sizeof(kiss_fft_cpx) * stwn);
            if (!rc) rc = (int)cudaMemcpy(d_stw, stw, sizeof(kiss_fft_cpx) * stwn, cudaMemcpyHostToDevice);
        }
        if (rc) {
            if (d_tw) cudaFree(d_tw);
            if (d_stw) cudaFree(d_stw);
            if (e) { free(e->h_tw); free(e); }
            pthread_mutex_unlock(&g_lock);
            return kf_cuda_fail(__FILE__, __LINE__, "device plan creation", rc);
        }
        memcpy(e->h_tw, cfg->twiddles, twbytes);
        e->device = dev;
        e->nfft = cfg->nfft;
        e->inverse = cfg->inverse;
        e->has_stw = (stw != NULL);
        e->plan.nfft = cfg->nfft;
        e->plan.inverse = cfg->inverse;
        e->plan.nstages = cfg->nstages;
        for (int s = 0; s < cfg->nstages; ++s) {
            e->plan.p[s] = cfg->factors[2 * s];
            e->plan.m[s] = cfg->factors[2 * s + 1];
        }
        e->plan.d_tw = d_tw;
        e->plan.d_stw = d_stw;
        e->plan.h_tw = e->h_tw;
        e->next = g_plans;
        g_plans = e;
    }
    pthread_mutex_unlock(&g_lock);
    *out = e;
    return 0;
}

static void kf_ctx_destroy(kf_ctx *c)
{
    cudaSetDevice(c->device);
    for (int i = 0; i < 3; ++i)
        if (c->d_buf[i]) cudaFree(c->d_buf[i]);
    for (int i = 0; i < 2; ++i)
        if (c->h_pin[i]) cudaFreeHost(c->h_pin[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    free(c);
}

void kiss_fft_cleanup(void)
{
    int cur = 0;
    cudaGetDevice(&cur);
    pthread_mutex_lock(&g_mp_lock);
    pthread_mutex_lock(&g_pool_lock);
    pthread_mutex_lock(&g_lock);
    for (kf_devplan *e = g_plans; e;) {
        kf_devplan *n = e->next;
        cudaSetDevice(e->device);
        cudaFree((void *)e->plan.d_tw);
        if (e->plan.d_stw) cudaFree((void *)e->plan.d_stw);
        for (int m = 0; m < 5; ++m)
            if (e->plan.d_gtw[m]) cudaFree(e->plan.d_gtw[m]);
        free(e->h_tw);
        free(e);
        e = n;
    }
    g_plans = NULL;
    for (int i = 0; i < KF_MP_SLOTS; ++i) {
        if (g_mp_bufs[i].ptr) {
            cudaSetDevice(g_mp_bufs[i].device);
            cudaFree(g_mp_bufs[i].ptr);
        }
        g_mp_bufs[i].ptr = NULL;
        g_mp_bufs[i].bytes = 0;
    }
    /* idle staging contexts go; a busy one belongs to a call that is still running (documented misuse) and is left */
    kf_ctx **pp = &g_pool;
    while (*pp) {
        kf_ctx *c = *pp;
        if (c->busy) {
            pp = &c->next;
        } else {
            *pp = c->next;
            kf_ctx_destroy(c);
        }
    }
    cudaSetDevice(cur);
    pthread_mutex_unlock(&g_lock);
    pthread_mutex_unlock(&g_pool_lock);
    pthread_mutex_unlock(&g_mp_lock);
}

/* work-buffer slot `i` of the long-row paths, at least `bytes` large, on the current device (caller holds g_mp_lock) */
static int kf_mp_scratch(int i, size_t bytes, void **out)
{
    int dev = 0;
    KF_CHECK(cudaGetDevice(&dev));
    kf_buf *b = &g_mp_bufs[i];
    if (b->ptr && (b->bytes < bytes || b->device != dev)) {
        cudaSetDevice(b->device);
        cudaFree(b->ptr);
        cudaSetDevice(dev);
        b->ptr = NULL;
        b->bytes = 0;
    }
    if (!b->ptr) {
        KF_CHECK(cudaMalloc(&b->ptr, bytes ? bytes : 1));
        b->bytes = bytes;
        b->device = dev;
    }
    *out = b->ptr;
    return 0;
}

static int kf_ctx_acquire(kf_ctx **out)
{
    int dev = 0;
    KF_CHECK(cudaGetDevice(&dev));
    pthread_mutex_lock(&g_pool_lock);
    kf_ctx *c;
    for (c = g_pool; c; c = c->next)
        if (!c->busy && c->device == dev) break;
    if (c) c->busy = 1;
    pthread_mutex_unlock(&g_pool_lock);
    if (!c) {
        c = (kf_ctx *)calloc(1, sizeof(*c));
        if (!c) return KISS_FFT_CUDA_ENOMEM;
        c->device = dev;
        c->busy = 1;
        int rc = (int)cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (rc) {
            free(c);
            return kf_cuda_fail(__FILE__, __LINE__, "cudaStreamCreate", rc);
        }
        pthread_mutex_lock(&g_pool_lock);
        c->next = g_pool;
        g_pool = c;
        pthread_mutex_unlock(&g_pool_lock);
    }
    *out = c;
    return 0;
}

static void kf_ctx_release(kf_ctx *c)
{
    if (!c) return;
    pthread_mutex_lock(&g_pool_lock);
    c->busy = 0;
    pthread_mutex_unlock(&g_pool_lock);
}

/* device scratch slot / pinned bounce buffer of a context, grown on demand (the context is owned by the caller) */
static int kf_ctx_dev(kf_ctx *c, int slot, size_t bytes, void **out)
{
    if (c->d_bytes[slot] < bytes || !c->d_buf[slot]) {
        if (c->d_buf[slot]) {
            KF_CHECK(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_buf[slot]);
            c->d_buf[slot] = NULL;
            c->d_bytes[slot] = 0;
        }
        KF_CHECK(cudaMalloc(&c->d_buf[slot], bytes ? bytes : 1));
        c->d_bytes[slot] = bytes;
    }
    *out = c->d_buf[slot];
    return 0;
}

static int kf_ctx_pin(kf_ctx *c, int slot, size_t bytes, void **out)
{
    if (c->h_bytes[slot] < bytes || !c->h_pin[slot]) {
        if (c->h_pin[slot]) {
            KF_CHECK(cudaStreamSynchronize(c->stream));
            cudaFreeHost(c->h_pin[slot]);
            c->h_pin[slot] = NULL;
            c->h_bytes[slot] = 0;
        }
        KF_CHECK(cudaHostAlloc(&c->h_pin[slot], bytes ? bytes : 1, cudaHostAllocDefault));
        c->h_bytes[slot] = bytes;
    }
    *out = c->h_pin[slot];
    return 0;
}

/* 0: pageable host memory (the driver would stage it synchronously), 1: pinned / registered host memory */
static int kf_is_pinned_host(const void *p)
{
    struct cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeHost;
}

static int kf_is_device_ptr(const void *p)
{
    struct cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

/* ---- execution: shared-memory kernels, or the multi-pass path for lengths that do not fit them --------------- */
/* One radix stage per launch, ping-ponging between two dense work buffers (slots 0, 1); the real modes add a work
 * buffer for the packed-complex array (slot 2) and a stand-alone split pass.  The work buffers are shared, so the
 * call holds g_mp_lock and waits for the stream before releasing it. */
static int kf_exec_multipass(int mode, const kf_devplan *dp, const void *d_in, void *d_out, long long howmany, long long in_dist,
                             long long out_dist, long long in_stride, void *stream)
{
    const kfcu_plan *pl = &dp->plan;
    const int N = pl->nfft, L = pl->nstages;
    const size_t dense = sizeof(kiss_fft_cpx) * (size_t)N * (size_t)howmany;
    pthread_mutex_lock(&g_mp_lock);
    void *w0 = NULL, *w1 = NULL, *wt = NULL;
    int rc = kf_mp_scratch(0, dense, &w0);
    if (!rc) rc = kf_mp_scratch(1, dense, &w1);
    if (!rc && (mode == KFCU_R2C || mode == KFCU_C2R)) rc = kf_mp_scratch(2, dense, &wt);
    const void *src = d_in;
    void *dst_final = d_out;
    long long sdist = in_dist, sstride = in_stride, ddist = out_dist;
    if (!rc && mode == KFCU_C2R) {          /* F -> T (dense), then the inverse complex transform of T */
        rc = kfcu_realpass(pl, 0, d_in, wt, howmany, in_dist, N, stream);
        src = wt;
        sdist = N;
        sstride = 1;
    }
    if (!rc && mode == KFCU_R2C) {          /* complex transform into T (dense), split pass afterwards */
        dst_final = wt;
        ddist = N;
    }
    const void *cur = src;
    for (int s = L - 1; !rc && s >= 0; --s) {
        const int first = (s == L - 1), last = (s == 0);
        void *out = last ? dst_final : ((cur == w0) ? w1 : w0);
        if (last && L == 1 && cur == dst_final) {   /* single stage, in place: go through a work buffer */
            rc = kfcu_stage(pl, s, cur, w0, howmany, sdist, N, sstride, first, 0, stream);
            if (!rc)
                rc = (int)cudaMemcpy2DAsync(dst_final, sizeof(kiss_fft_cpx) * (size_t)ddist, w0, sizeof(kiss_fft_cpx) * (size_t)N,
                                            sizeof(kiss_fft_cpx) * (size_t)N, (size_t)howmany, cudaMemcpyDeviceToDevice,
                                            (cudaStream_t)stream);
            break;
        }
        rc = kfcu_stage(pl, s, cur, out, howmany, first ? sdist : N, last ? ddist : N, first ? sstride : 1, first, last, stream);
        cur = out;
    }
    if (!rc && mode == KFCU_R2C) rc = kfcu_realpass(pl, 1, wt, d_out, howmany, N, out_dist, stream);
    if (!rc) rc = (int)cudaStreamSynchronize((cudaStream_t)stream);
    pthread_mutex_unlock(&g_mp_lock);
    return rc;
}

#ifndef FIXED_POINT
/* Four-step path for long contiguous rows (float / double): N = N1*N2 with a fused column plan for both factors.
 *   x[n1*N2 + n2]  --columns of length N1, times W_N^(n2*k1)-->  A[n2*N1 + k1]  --columns of length N2-->  X[k2*N1 + k1]
 * Two launches and four passes over the data instead of the 2*L of the stage-per-launch path.  A different
 * factorisation than kf_work's, so the result matches the reference to rounding only (hence not for fixed point).
 * Default for long float/double rows since round 2 (2.5-3x the stage-per-launch path on B200,
 * profiles/r02/long_rows_*_{fourstep,multipass}.jsonl); KISSFFT_FOURSTEP=0 selects the stage-per-launch path. */
static int kf_fourstep_split(int nfft, int *n1, int *n2)
{
    int best = 0;
    for (int a = 2; (long long)a * a <= (long long)nfft; ++a) {
        if (nfft % a) continue;
        const int b = nfft / a;
        if (kfcu_has_fourstep(a) && kfcu_has_fourstep(b)) best = a;     /* the largest a <= sqrt(nfft): most balanced split */
    }
    if (!best) return 0;
    *n1 = best;
    *n2 = nfft / best;
    return 1;
}

static int kf_exec_fourstep(int mode, const kf_devplan *dp, int n1, int n2, const void *d_in, void *d_out, long long howmany,
                            long long in_dist, long long out_dist, void *stream)
{
    const int N = dp->plan.nfft, inverse = dp->plan.inverse;
    kiss_fft_cfg c1 = kiss_fft_alloc(n1, inverse, NULL, NULL), c2 = kiss_fft_alloc(n2, inverse, NULL, NULL);
    const kf_devplan *p1 = NULL, *p2 = NULL;
    int rc = (c1 && c2) ? 0 : KISS_FFT_CUDA_ENOMEM;
    if (!rc) rc = kf_get_devplan(c1, NULL, &p1);
    if (!rc) rc = kf_get_devplan(c2, NULL, &p2);
    free(c1);
    free(c2);
    if (rc) return rc;
    const size_t dense = sizeof(kiss_fft_cpx) * (size_t)N * (size_t)howmany;
    pthread_mutex_lock(&g_mp_lock);
    void *work = NULL, *wt = NULL;
    rc = kf_mp_scratch(0, dense, &work);
    /* real transforms of a long row: the packed complex transform runs four-step on a dense array T[] (slot 2) and the
     * split pass (kiss_fftr.c:88-116 / :131-153) is a stand-alone kernel before / after it */
    if (!rc && mode != KFCU_C2C) rc = kf_mp_scratch(2, dense, &wt);
    const void *src = d_in;
    void *dst = d_out;
    if (!rc && mode == KFCU_C2R) {
        rc = kfcu_realpass(&dp->plan, 0, d_in, wt, howmany, in_dist, N, stream);
        src = wt;
    }
    if (mode == KFCU_R2C) dst = wt;
    if (!rc) rc = kfcu_exec_fourstep((kfcu_plan *)&p1->plan, 0, src, work, howmany, n2, dp->plan.d_tw, stream);
    if (!rc) rc = kfcu_exec_fourstep((kfcu_plan *)&p2->plan, 1, work, dst, howmany, n1, NULL, stream);
    if (!rc && mode == KFCU_R2C) rc = kfcu_realpass(&dp->plan, 1, wt, d_out, howmany, N, out_dist, stream);
    if (!rc) rc = (int)cudaStreamSynchronize((cudaStream_t)stream);      /* the work buffers are shared */
    pthread_mutex_unlock(&g_mp_lock);
    return rc;
}
#endif

static int kf_exec(int mode, const kf_devplan *dp, const void *d_in, void *d_out, long long howmany, long long in_dist,
                   long long out_dist, long long in_stride, void *stream)
{
    int rc = kfcu_exec(mode, (kfcu_plan *)&dp->plan, d_in, d_out, howmany, in_dist, out_dist, in_stride, stream);
    if (rc != KFCU_ETOOBIG) return rc;
#ifndef FIXED_POINT
    const char *opt = getenv("KISSFFT_FOURSTEP");
    int n1 = 0, n2 = 0;
    const int N = dp->plan.nfft;
    /* complex rows must be dense (the passes view each row as an N1 x N2 array); real rows may be any distance apart
     * because their split pass reads / writes them and the complex transform runs on the dense T[] */
    const int dense_ok = mode == KFCU_C2C ? (in_stride == 1 && in_dist == N && out_dist == N)
                                          : (mode == KFCU_R2C ? in_dist == N : out_dist == N);
    if (!(opt && opt[0] == '0') && (mode == KFCU_C2C || mode == KFCU_R2C || mode == KFCU_C2R) && dense_ok &&
        kf_fourstep_split(N, &n1, &n2))
        return kf_exec_fourstep(mode, dp, n1, n2, d_in, d_out, howmany, in_dist, out_dist, stream);
#endif
    return kf_exec_multipass(mode, dp, d_in, d_out, howmany, in_dist, out_dist, in_stride, stream);
}

/* ---- device-pointer batched entry points ---------------------------------------------------------------- */

int kiss_fft_batch_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t howmany, size_t in_dist,
                       size_t out_dist, int in_stride, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_1D || !d_in || !d_out || in_stride < 1) {
        KF_ERROR("kiss_fft_batch_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg, NULL, &dp));
    KF_CHECK(kf_exec(KFCU_C2C, dp, d_in, d_out, (long long)howmany, (long long)in_dist, (long long)out_dist,
                       (long long)in_stride, stream));
    return 0;
}

int kiss_fft_axis_pass_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t ncols, size_t col_stride,
                           void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_1D || !d_in || !d_out || col_stride < 1) {
        KF_ERROR("kiss_fft_axis_pass_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg, NULL, &dp));
    /* column i: elements d_in[i + j*col_stride]; written as row i: d_out[i*nfft + k]  (kiss_fftnd.c:176-177) */
    const int mode = (col_stride == 1 && ncols == 1) ? KFCU_C2C : KFCU_C2C_COL;
    KF_CHECK(kf_exec(mode, dp, d_in, d_out, (long long)ncols, 1, (long long)cfg->nfft, (long long)col_stride, stream));
    return 0;
}

int kiss_fft_planes_pass_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t nplanes, size_t ncols,
                             size_t col_stride, size_t in_plane_dist, size_t out_plane_dist, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_1D || !d_in || !d_out || col_stride < 1) {
        KF_ERROR("kiss_fft_planes_pass_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg, NULL, &dp));
    KF_CHECK(kfcu_exec_planes((kfcu_plan *)&dp->plan, d_in, d_out, (long long)nplanes, (long long)ncols, (long long)col_stride,
                              (long long)in_plane_dist, (long long)out_plane_dist, stream));
    return 0;
}

int kiss_fft_planes_pass_peers2_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *const *d_peers, int npeers,
                                    size_t nplanes, size_t cols_per_peer, size_t peer_col_dist, size_t col_stride,
                                    size_t in_plane_dist, size_t out_plane_dist, size_t out_col_dist, int max_ctas, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_1D || !d_in || !d_peers || npeers < 1 || npeers > 16 || col_stride < 1 ||
        peer_col_dist < cols_per_peer) {
        KF_ERROR("kiss_fft_planes_pass_peers_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg, NULL, &dp));
    KF_CHECK(kfcu_exec_planes_peers((kfcu_plan *)&dp->plan, d_in, (void *const *)d_peers, npeers, (long long)nplanes,
                                    (long long)cols_per_peer, (long long)peer_col_dist, (long long)col_stride,
                                    (long long)in_plane_dist, (long long)out_plane_dist, (long long)out_col_dist, max_ctas, stream));
    return 0;
}

int kiss_fft_planes_pass_peers_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *const *d_peers, int npeers,
                                   size_t nplanes, size_t cols_per_peer, size_t col_stride, size_t in_plane_dist,
                                   size_t out_plane_dist, void *stream)
{
    return kiss_fft_planes_pass_peers2_dev(cfg, d_in, d_peers, npeers, nplanes, cols_per_peer, cols_per_peer, col_stride, in_plane_dist,
                                           out_plane_dist, 0, 0, stream);
}

static int kf_real_args_ok(const void *d_real, size_t real_dist)
{
    return ((uintptr_t)d_real % (2 * sizeof(kiss_fft_scalar))) == 0 && (real_dist % 2) == 0;
}

int kiss_fftr_batch_dev(kiss_fftr_cfg cfg, const kiss_fft_scalar *d_time, kiss_fft_cpx *d_freq, size_t howmany,
                        size_t time_dist, size_t freq_dist, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_R || !d_time || !d_freq) {
        KF_ERROR("kiss_fftr_batch_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    if (cfg->substate->inverse) {
        KF_ERROR("kiss fft usage error: improper alloc"); /* kiss_fftr.c:69-72 */
        return KISS_FFT_CUDA_EINVAL;
    }
    if (!kf_real_args_ok(d_time, time_dist)) {
        KF_ERROR("kiss_fftr_batch_dev: real rows must be 2*sizeof(scalar) aligned and an even distance apart");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg->substate, cfg->super_twiddles, &dp));
    KF_CHECK(kf_exec(KFCU_R2C, dp, d_time, d_freq, (long long)howmany, (long long)(time_dist / 2),
                       (long long)freq_dist, 1, stream));
    return 0;
}

int kiss_fftri_batch_dev(kiss_fftr_cfg cfg, const kiss_fft_cpx *d_freq, kiss_fft_scalar *d_time, size_t howmany,
                         size_t freq_dist, size_t time_dist, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_R || !d_time || !d_freq) {
        KF_ERROR("kiss_fftri_batch_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    if (cfg->substate->inverse == 0) {
        KF_ERROR("kiss fft usage error: improper alloc"); /* kiss_fftr.c:124-127 */
        return KISS_FFT_CUDA_EINVAL;
    }
    if (!kf_real_args_ok(d_time, time_dist)) {
        KF_ERROR("kiss_fftri_batch_dev: real rows must be 2*sizeof(scalar) aligned and an even distance apart");
        return KISS_FFT_CUDA_EINVAL;
    }
    const kf_devplan *dp;
    KF_CHECK(kf_get_devplan(cfg->substate, cfg->super_twiddles, &dp));
    KF_CHECK(kf_exec(KFCU_C2R, dp, d_freq, d_time, (long long)howmany, (long long)freq_dist,
                       (long long)(time_dist / 2), 1, stream));
    return 0;
}

/* kiss_fftnd.c:156-188 on device buffers.  Pass k views the current buffer as dims[k] x stride, transforms
 * every column and stores it as a row, ping-ponging so that the last pass lands in d_out. */
static int kf_fftnd_dev_locked(kiss_fftnd_cfg st, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, kiss_fft_cpx *d_work,
                               void *stream)
{
    const size_t bytes = sizeof(kiss_fft_cpx) * (size_t)st->dimprod;
    const kiss_fft_cpx *bufin = d_in;
    kiss_fft_cpx *bufout;
    if (st->ndims == 0) return 0;
    if (st->ndims & 1) {
        bufout = d_out;
        if (d_in == d_out) {
            KF_CHECK(cudaMemcpyAsync(d_work, d_in, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            bufin = d_work;
        }
    } else {
        bufout = d_work;
    }
    for (int k = 0; k < st->ndims; ++k) {
        const int curdim = st->dims[k];
        const long long stride = st->dimprod / curdim;
        KF_CHECK(kiss_fft_axis_pass_dev(st->states[k], bufin, bufout, (size_t)stride, (size_t)stride, stream));
        if (bufout == d_work) {
            bufout = d_out;
            bufin = d_work;
        } else {
            bufout = d_work;
            bufin = d_out;
        }
    }
    return 0;
}

/* In-layout variant (default whenever every leading axis has a fused column plan; KISSFFT_FFTND_INLAYOUT=0 selects the
 * transposing sweeps): every axis is transformed where it lies -- strided columns in, the same strided columns out --
 * instead of kiss_fftnd.c's transposing sweeps, and the last axis is a plain row pass.  Axes still run in the
 * reference's order 0,1,... on the same operands, so the results (fixed point included) are the same bits; no work
 * buffer is needed (the reference's tmpbuf is 8 GiB at 1024^3, kiss_fftnd.c:38). */
static int kf_inlayout_enabled(void)
{
    const char *opt = getenv("KISSFFT_FFTND_INLAYOUT");
    return !(opt && opt[0] == '0');
}

/* can the leading `naxes` axes of `st` be transformed where they lie, with `inner` contiguous elements below the last
 * of them (inner == 1: the last axis is a plain row pass)? */
static int kf_axes_inlayout_ok(kiss_fftnd_cfg st, int naxes, long long inner)
{
    if (!kf_inlayout_enabled()) return 0;
    for (int k = 0; k < naxes; ++k) {
        const int is_rows = (k == naxes - 1 && inner == 1);
        if (!is_rows && !kfcu_has_colcol(st->dims[k])) return 0;
    }
    return 1;
}

/* axes 0..naxes-1 of the array [dims[0]]..[dims[naxes-1]][inner], in the reference's order, each where it lies: the
 * first pass reads d_in and writes d_out, the others run in place on d_out */
static int kf_axes_inlayout(kiss_fftnd_cfg st, int naxes, long long inner, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out,
                            void *stream)
{
    long long nplanes = 1, below = inner;
    for (int k = 0; k < naxes; ++k) below *= st->dims[k];
    for (int k = 0; k < naxes; ++k) {
        const int n = st->dims[k];
        below /= n;                                   /* elements per index step of axis k */
        const kiss_fft_cpx *src = (k == 0) ? d_in : d_out;
        const kf_devplan *dp;
        KF_CHECK(kf_get_devplan(st->states[k], NULL, &dp));
        if (below == 1)
            KF_CHECK(kf_exec(KFCU_C2C, dp, src, d_out, nplanes, n, n, 1, stream));
        else
            KF_CHECK(kfcu_exec_fourstep((kfcu_plan *)&dp->plan, 1, src, d_out, nplanes, below, NULL, stream));
        nplanes *= n;
    }
    return 0;
}

#ifndef FIXED_POINT
/* 3-D, float / double: three transposing passes, none of which strides over whole planes.
 *
 * kiss_fftnd.c:156-188 views the buffer as dims[k] x stride in every sweep, so each sweep walks columns whose elements lie
 * a whole plane apart (8 MiB at 1024^3): every row segment of a tile sits on its own page and the pass is bound by
 * address translation, not by HBM (profiles/r02/stride_experiment.txt: time ~ 1 / segment width, whatever the stride's
 * alignment).  A transposing pass may put each finished row wherever it likes at no cost -- rows are written as whole
 * contiguous lines anyway -- which is enough to keep the columns of EVERY pass inside one plane (row pitch = one line):
 *     in  [i0][i1][i2]  --axis 1 of plane i0-->  A [i2][i0][k1]     (row (i0, i2) of the result goes to A[i2][i0])
 *     A   [i2][i0][k1]  --axis 0 of plane i2-->  B [k1][i2][k0]
 *     B   [k1][i2][k0]  --axis 2 of plane k1-->  out [k0][k1][k2]   natural order, as kiss_fftnd leaves it
 * Axis order 1, 0, 2 instead of the reference's 0, 1, 2: equal to rounding in float / double, NOT the same bits in fixed
 * point, which therefore keeps the in-layout path.  Needs one work buffer the size of the array (A lives in d_out, B in
 * d_work) -- what the reference carries in every kiss_fftnd cfg (tmpbuf, kiss_fftnd.c:38). */
static int kf_fftnd3_permuted(kiss_fftnd_cfg st, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, kiss_fft_cpx *d_work, void *stream)
{
    const long long d0 = st->dims[0], d1 = st->dims[1], d2 = st->dims[2];
    const kf_devplan *p0, *p1, *p2;
    KF_CHECK(kf_get_devplan(st->states[0], NULL, &p0));
    KF_CHECK(kf_get_devplan(st->states[1], NULL, &p1));
    KF_CHECK(kf_get_devplan(st->states[2], NULL, &p2));
    void *dst = d_out;     /* (plan, in, {out}, 1, planes, columns, columns, column stride, plane pitch in, plane pitch out, row pitch out) */
    KF_CHECK(kfcu_exec_planes_peers((kfcu_plan *)&p1->plan, d_in, &dst, 1, d0, d2, d2, d2, d1 * d2, d1, d0 * d1, 0, stream));
    dst = d_work;
    KF_CHECK(kfcu_exec_planes_peers((kfcu_plan *)&p0->plan, d_out, &dst, 1, d2, d1, d1, d1, d0 * d1, d0, d2 * d0, 0, stream));
    dst = d_out;
    KF_CHECK(kfcu_exec_planes_peers((kfcu_plan *)&p2->plan, d_work, &dst, 1, d1, d0, d0, d0, d2 * d0, d2, d1 * d2, 0, stream));
    return 0;
}

/* auto: when a sweep of the plain layout would stride further than the tensor-map input ring accepts (kf_tmap.h: 128 KiB)
 * and every axis length has a ring variant of the transposing pass -- measured on B200 (profiles/r02/fftnd_strategies.txt):
 * 1024^3 11.0 -> 9.02 ms; 512^3, whose passes have no ring variant yet, 1.48 -> 1.70 ms, so it stays in layout.
 * KISSFFT_FFTND_PERMUTE=0 / 1 forces the choice. */
static int kf_fftnd3_permuted_wanted(kiss_fftnd_cfg st, const kiss_fft_cpx *d_in, const kiss_fft_cpx *d_out)
{
    if (st->ndims != 3 || d_in == d_out) return 0;
    const char *opt = getenv("KISSFFT_FFTND_PERMUTE");
    if (opt && opt[0] == '0') return 0;
    if (opt && opt[0] == '1') return 1;
    const size_t plane = sizeof(kiss_fft_cpx) * (size_t)st->dims[1] * (size_t)st->dims[2];
    return plane > ((size_t)128 << 10) && kfcu_has_colring(st->dims[0]) && kfcu_has_colring(st->dims[1]) && kfcu_has_colring(st->dims[2]);
}
#endif

int kiss_fftnd_dev(kiss_fftnd_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, kiss_fft_cpx *d_work, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_ND || !d_in || !d_out) {
        KF_ERROR("kiss_fftnd_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
#ifndef FIXED_POINT
    if (kf_fftnd3_permuted_wanted(cfg, d_in, d_out)) {
        if (d_work) return kf_fftnd3_permuted(cfg, d_in, d_out, d_work, stream);
        /* internal work buffer from a staging context (kept for the next call); if the device cannot spare it the
         * in-layout path below needs none */
        kf_ctx *cx = NULL;
        void *w = NULL;
        if (kf_ctx_acquire(&cx) == 0) {
            if (kf_ctx_dev(cx, 2, sizeof(kiss_fft_cpx) * (size_t)cfg->dimprod, &w) == 0) {
                int rc = kf_fftnd3_permuted(cfg, d_in, d_out, (kiss_fft_cpx *)w, stream);
                int e = (int)cudaStreamSynchronize((cudaStream_t)stream);      /* the context goes back to the pool */
                kf_ctx_release(cx);
                return rc ? rc : e;
            }
            cudaGetLastError();
            kf_ctx_release(cx);
        }
    }
#endif
    if (cfg->ndims >= 2 && kf_axes_inlayout_ok(cfg, cfg->ndims, 1)) return kf_axes_inlayout(cfg, cfg->ndims, 1, d_in, d_out, stream);
    if (d_work) return kf_fftnd_dev_locked(cfg, d_in, d_out, d_work, stream);
    /* internal scratch from a staging context: this variant waits for completion before giving the context back */
    kf_ctx *cx = NULL;
    KF_CHECK(kf_ctx_acquire(&cx));
    void *w = NULL;
    int rc = kf_ctx_dev(cx, 2, sizeof(kiss_fft_cpx) * (size_t)cfg->dimprod, &w);
    if (!rc) rc = kf_fftnd_dev_locked(cfg, d_in, d_out, (kiss_fft_cpx *)w, stream);
    int e = (int)cudaStreamSynchronize((cudaStream_t)stream);
    kf_ctx_release(cx);
    return rc ? rc : e;
}

/* kiss_fftndr.c:86-110 and :112-132 on device buffers.
 *
 * The reference gathers every frequency bin into its own contiguous array, runs kiss_fftnd on it and scatters
 * the result back.  Here the bins stay interleaved: with B = dimReal/2+1 the half spectra form the array
 * [d0][d1]..[d(n-2)][B]; a transposing axis pass (kiss_fftnd.c:172-178) over the leading axis turns it into
 * [d1]..[d(n-2)][B][d0], and after all n-1 leading axes it is [B][d0]..[d(n-2)], i.e. exactly the reference's
 * bin-major tmp2 (kiss_fftndr.c:101-102) already transformed.  One transpose brings it to [dimOther][B].
 * Every 1-D transform sees the reference's operands and the axes run in the reference's order 0,1,..., so the
 * fixed-point results are bit-identical. */
static int kf_leading_axes_dev(kiss_fftnd_cfg nd, size_t total, kiss_fft_cpx **a, kiss_fft_cpx **b, void *stream)
{
    for (int k = 0; k < nd->ndims; ++k) {
        const size_t cols = total / (size_t)nd->dims[k];
        KF_CHECK(kiss_fft_axis_pass_dev(nd->states[k], *a, *b, cols, cols, stream));
        kiss_fft_cpx *t = *a;
        *a = *b;
        *b = t;
    }
    return 0;
}

int kiss_fftndr_dev(kiss_fftndr_cfg cfg, const kiss_fft_scalar *d_time, kiss_fft_cpx *d_freq, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_NDR || !d_time || !d_freq) {
        KF_ERROR("kiss_fftndr_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    if (cfg->inverse) {
        KF_ERROR("kiss fft usage error: improper alloc");
        return KISS_FFT_CUDA_EINVAL;
    }
    const size_t nrbins = (size_t)cfg->dimReal / 2 + 1;
    const size_t total = (size_t)cfg->dimOther * nrbins;
    if (cfg->ndims == 1) return kiss_fftr_batch_dev(cfg->cfg_r, d_time, d_freq, 1, (size_t)cfg->dimReal, nrbins, stream);
    if (kf_axes_inlayout_ok(cfg->cfg_nd, cfg->cfg_nd->ndims, (long long)nrbins)) {
        /* no gather / scatter at all (kiss_fftndr.c:101-108 folded away): the half spectra land in the caller's array
         * [dimOther][B] and every leading axis is transformed where it lies -- stream-ordered, no scratch, no sync */
        KF_CHECK(kiss_fftr_batch_dev(cfg->cfg_r, d_time, d_freq, (size_t)cfg->dimOther, (size_t)cfg->dimReal, nrbins, stream));
        return kf_axes_inlayout(cfg->cfg_nd, cfg->cfg_nd->ndims, (long long)nrbins, d_freq, d_freq, stream);
    }
    kf_ctx *cx = NULL;
    KF_CHECK(kf_ctx_acquire(&cx));
    void *w0 = NULL, *w1 = NULL;
    int rc = kf_ctx_dev(cx, 0, sizeof(kiss_fft_cpx) * total, &w0);
    if (!rc) rc = kf_ctx_dev(cx, 1, sizeof(kiss_fft_cpx) * total, &w1);
    kiss_fft_cpx *a = (kiss_fft_cpx *)w0, *b = (kiss_fft_cpx *)w1;
    if (!rc) rc = kiss_fftr_batch_dev(cfg->cfg_r, d_time, a, (size_t)cfg->dimOther, (size_t)cfg->dimReal, nrbins, stream);
    if (!rc) rc = kf_leading_axes_dev(cfg->cfg_nd, total, &a, &b, stream);
    /* a: [B][dimOther] -> d_freq: [dimOther][B] */
    if (!rc) rc = kfcu_transpose(a, d_freq, (long long)nrbins, (long long)cfg->dimOther, stream);
    int e = (int)cudaStreamSynchronize((cudaStream_t)stream);       /* the scratch goes back to the pool */
    kf_ctx_release(cx);
    if (!rc) rc = e;
    if (rc) return kf_cuda_fail(__FILE__, __LINE__, "kiss_fftndr_dev", rc);
    return 0;
}

int kiss_fftndri_dev(kiss_fftndr_cfg cfg, const kiss_fft_cpx *d_freq, kiss_fft_scalar *d_time, void *stream)
{
    if (!cfg || cfg->magic != KF_MAGIC_NDR || !d_time || !d_freq) {
        KF_ERROR("kiss_fftndri_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    if (!cfg->inverse) {
        KF_ERROR("kiss fft usage error: improper alloc");
        return KISS_FFT_CUDA_EINVAL;
    }
    const size_t nrbins = (size_t)cfg->dimReal / 2 + 1;
    const size_t total = (size_t)cfg->dimOther * nrbins;
    if (cfg->ndims == 1) return kiss_fftri_batch_dev(cfg->cfg_r, d_freq, d_time, 1, nrbins, (size_t)cfg->dimReal, stream);
    kf_ctx *cx = NULL;
    KF_CHECK(kf_ctx_acquire(&cx));
    void *w0 = NULL, *w1 = NULL;
    int rc = kf_ctx_dev(cx, 0, sizeof(kiss_fft_cpx) * total, &w0);
    kiss_fft_cpx *a = (kiss_fft_cpx *)w0, *b = NULL;
    kiss_fftnd_cfg nd = cfg->cfg_nd;
    if (!rc && kf_axes_inlayout_ok(nd, nd->ndims, (long long)nrbins)) {
        /* leading axes where they lie: the caller's spectrum is read once (it is const), the rest runs in place in `a` */
        rc = kf_axes_inlayout(nd, nd->ndims, (long long)nrbins, d_freq, a, stream);
        if (!rc) rc = kiss_fftri_batch_dev(cfg->cfg_r, a, d_time, (size_t)cfg->dimOther, nrbins, (size_t)cfg->dimReal, stream);
    } else {
        if (!rc) rc = kf_ctx_dev(cx, 1, sizeof(kiss_fft_cpx) * total, &w1);
        b = (kiss_fft_cpx *)w1;
        /* first leading-axis pass reads the caller's buffer directly ([d0].. [B] has d0 leading already) */
        if (!rc) {
            const size_t cols = total / (size_t)nd->dims[0];
            rc = kiss_fft_axis_pass_dev(nd->states[0], d_freq, a, cols, cols, stream);
        }
        for (int k = 1; !rc && k < nd->ndims; ++k) {
            const size_t cols = total / (size_t)nd->dims[k];
            rc = kiss_fft_axis_pass_dev(nd->states[k], a, b, cols, cols, stream);
            kiss_fft_cpx *t = a; a = b; b = t;
        }
        /* a: [B][dimOther] -> b: [dimOther][B], then the real inverse of every row (kiss_fftndr.c:127-131) */
        if (!rc) rc = kfcu_transpose(a, b, (long long)nrbins, (long long)cfg->dimOther, stream);
        if (!rc) rc = kiss_fftri_batch_dev(cfg->cfg_r, b, d_time, (size_t)cfg->dimOther, nrbins, (size_t)cfg->dimReal, stream);
    }
    int e = (int)cudaStreamSynchronize((cudaStream_t)stream);
    kf_ctx_release(cx);
    if (!rc) rc = e;
    if (rc) return kf_cuda_fail(__FILE__, __LINE__, "kiss_fftndri_dev", rc);
    return 0;
}

/* ---- host-pointer batched entry points: chunked H2D / kernel / D2H pipeline over 3 streams ---------------- */
typedef int (*kf_chunk_fn)(void *cfg, const void *d_in, void *d_out, size_t howmany, void *stream);

static int kf_chunk_c2c(void *cfg, const void *d_in, void *d_out, size_t n, void *stream)
{
    kiss_fft_cfg c = (kiss_fft_cfg)cfg;
    return kiss_fft_batch_dev(c, (const kiss_fft_cpx *)d_in, (kiss_fft_cpx *)d_out, n, (size_t)c->nfft, (size_t)c->nfft, 1, stream);
}
static int kf_chunk_r2c(void *cfg, const void *d_in, void *d_out, size_t n, void *stream)
{
    kiss_fftr_cfg c = (kiss_fftr_cfg)cfg;
    return kiss_fftr_batch_dev(c, (const kiss_fft_scalar *)d_in, (kiss_fft_cpx *)d_out, n, (size_t)c->nfft,
                               (size_t)c->nfft / 2 + 1, stream);
}
static int kf_chunk_c2r(void *cfg, const void *d_in, void *d_out, size_t n, void *stream)
{
    kiss_fftr_cfg c = (kiss_fftr_cfg)cfg;
    return kiss_fftri_batch_dev(c, (const kiss_fft_cpx *)d_in, (kiss_fft_scalar *)d_out, n, (size_t)c->nfft / 2 + 1,
                                (size_t)c->nfft, stream);
}

/* One host-pointer batched call = `nlanes` lanes working through the chunks of the batch round-robin.  A lane owns a
 * staging context (stream, device chunk buffers, pinned bounce buffers) and runs, per chunk,
 *     [memcpy caller -> pinned]  H2D  kernel  D2H  [memcpy pinned -> caller]
 * in order on its own host thread; the lanes overlap each other, so the copy engines (both directions), the SMs and
 * the host cores doing the bounce copies are all busy at once.  The bracketed steps exist only for PAGEABLE caller
 * memory -- what every drop-in caller of the reference passes (test/benchkiss.c:76-79, tools/fftutil.c:24-27): a
 * cudaMemcpyAsync on it would be staged synchronously by the driver, one direction at a time.  Pinned caller buffers
 * are handed to the copy engines directly.  KISSFFT_HOST_LANES / KISSFFT_CHUNK_MIB override the defaults. */
/* defaults from the sweep in profiles/r02/e2e_probe_lanes_chunks.jsonl (16-core host, PCIe Gen5): pinned caller buffers are
 * PCIe-bound from 2 lanes x 16 MiB on; pageable ones are bound by the host's bounce copies and want many lanes of
 * small chunks */
#define KF_CHUNK_MIB_PINNED 32
#define KF_CHUNK_MIB_PAGEABLE 4
#define KF_MAX_LANES 16

typedef struct { size_t first, n; } kf_span;   /* rows [first, first + n) of the batch */

typedef struct {
    kf_chunk_fn fn;
    void *cfg;
    const char *in;
    char *out;
    size_t rows, in_row_bytes, out_row_bytes;   /* rows = the largest chunk (buffer size) */
    const kf_span *chunks;                      /* the chunks of the batch, in order; lanes take the next one when free */
    size_t nchunks, *next;
    int lane, nlanes, device, in_pinned, out_pinned;
    kf_ctx *cx;
    int rc;
} kf_lane;

static void *kf_lane_main(void *arg)
{
    kf_lane *L = (kf_lane *)arg;
    int rc = (int)cudaSetDevice(L->device);
    kf_ctx *cx = L->cx;
    void *din = NULL, *dout = NULL, *pin = NULL, *pout = NULL;
    if (!rc) rc = kf_ctx_dev(cx, 0, L->rows * L->in_row_bytes, &din);
    if (!rc) rc = kf_ctx_dev(cx, 1, L->rows * L->out_row_bytes, &dout);
    if (!rc && !L->in_pinned) rc = kf_ctx_pin(cx, 0, L->rows * L->in_row_bytes, &pin);
    if (!rc && !L->out_pinned) rc = kf_ctx_pin(cx, 1, L->rows * L->out_row_bytes, &pout);
    while (!rc) {
        const size_t c = __atomic_fetch_add(L->next, 1, __ATOMIC_RELAXED);
        if (c >= L->nchunks) break;
        const size_t first = L->chunks[c].first, n = L->chunks[c].n;
        const void *src = L->in + first * L->in_row_bytes;
        void *dst = L->out + first * L->out_row_bytes;
        if (!L->in_pinned) {
            memcpy(pin, src, n * L->in_row_bytes);
            src = pin;
        }
        rc = (int)cudaMemcpyAsync(din, src, n * L->in_row_bytes, cudaMemcpyHostToDevice, cx->stream);
        if (!rc) rc = L->fn(L->cfg, din, dout, n, cx->stream);
        if (!rc) rc = (int)cudaMemcpyAsync(L->out_pinned ? dst : pout, dout, n * L->out_row_bytes, cudaMemcpyDeviceToHost, cx->stream);
        const int e = (int)cudaStreamSynchronize(cx->stream);
        if (!rc) rc = e;
        if (!rc && !L->out_pinned) memcpy(dst, pout, n * L->out_row_bytes);
    }
    L->rc = rc;
    return NULL;
}

/* Cuts the batch into chunks of `rows` rows with a ramp at both ends (rows/8, rows/8, rows/4, rows/2, rows ... rows,
 * rows/2, rows/4, rows/8, rows/8): nothing leaves the device before the first chunk has gone in and come back, and nothing
 * enters while the last one comes out, so the first and the last chunk are the part of the call where only one PCIe
 * direction is busy -- they should be small, the chunks in between large (profiles/r02/pcie_probe.jsonl: both directions
 * together carry 2 x 50 GB/s, a call with uniform 32 MiB chunks reached 2 x 44).  All sizes are multiples of 64 rows. */
static size_t kf_make_chunks(size_t howmany, size_t rows, int ramp, kf_span **out)
{
    size_t cap = howmany / rows + 12, n = 0, pos = 0;
    kf_span *v = (kf_span *)malloc(sizeof(kf_span) * cap);
    if (!v) return 0;
    size_t head[4], nh = 0;
    if (ramp && rows >= 512 && howmany >= 6 * rows) {
        const size_t div[4] = {8, 8, 4, 2};
        for (int i = 0; i < 4; ++i) head[nh++] = (rows / div[i]) & ~(size_t)63;
    }
    size_t tail_rows = 0;
    for (size_t i = 0; i < nh; ++i) tail_rows += head[i];
    for (size_t i = 0; i < nh; ++i) { v[n].first = pos; v[n].n = head[i]; pos += head[i]; ++n; }
    while (howmany - pos > tail_rows + rows) { v[n].first = pos; v[n].n = rows; pos += rows; ++n; }
    if (howmany - pos > tail_rows) { v[n].first = pos; v[n].n = howmany - pos - tail_rows; pos += v[n].n; ++n; }
    for (size_t i = nh; i-- > 0;) { v[n].first = pos; v[n].n = head[i]; pos += head[i]; ++n; }
    if (pos < howmany) { v[n].first = pos; v[n].n = howmany - pos; ++n; }     /* no ramp: the remainder */
    *out = v;
    return n;
}

/* testing aid (host logic only, no CUDA call): the chunk list kf_host_pipeline would use */
size_t kiss_fft_cuda_debug_chunks(size_t howmany, size_t rows, int ramp, size_t *first, size_t *n, size_t cap)
{
    kf_span *v = NULL;
    const size_t k = kf_make_chunks(howmany, rows, ramp, &v);
    for (size_t i = 0; i < k && i < cap; ++i) { first[i] = v[i].first; n[i] = v[i].n; }
    free(v);
    return k;
}

static int kf_host_lanes(size_t nchunks, int pinned)
{
    int lanes = 4;
    if (!pinned) {
        /* bounce copies run on the host cores: three quarters of them, shared with the other ranks of this node */
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        const char *lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = (lw && atoi(lw) > 0) ? atoi(lw) : 1;
        lanes = (int)(ncpu * 3 / 4 / ranks);
        if (lanes < 2) lanes = 2;
        if (lanes > 12) lanes = 12;
    }
    const char *env = getenv("KISSFFT_HOST_LANES");
    if (env && atoi(env) > 0) lanes = atoi(env);
    if (lanes > KF_MAX_LANES) lanes = KF_MAX_LANES;
    if ((size_t)lanes > nchunks) lanes = (int)nchunks;
    return lanes < 1 ? 1 : lanes;
}

static int kf_host_pipeline(kf_chunk_fn fn, void *cfg, const void *in, void *out, size_t howmany, size_t in_row_bytes,
                            size_t out_row_bytes)
{
    if (howmany == 0) return 0;
    const int in_pinned = kf_is_pinned_host(in), out_pinned = kf_is_pinned_host(out);
    size_t chunk_mib = (in_pinned && out_pinned) ? KF_CHUNK_MIB_PINNED : KF_CHUNK_MIB_PAGEABLE;
    const char *env = getenv("KISSFFT_CHUNK_MIB");
    if (env && atoi(env) > 0 && atoi(env) <= 1024) chunk_mib = (size_t)atoi(env);
    size_t rows = (chunk_mib << 20) / (in_row_bytes > out_row_bytes ? in_row_bytes : out_row_bytes);
    if (rows < 1) rows = 1;
    if (rows >= 64) rows &= ~(size_t)63; /* whole tiles for every fused plan (tpc <= 16) and 16-byte aligned chunk sizes */
    if (rows > howmany) rows = howmany;
    int dev = 0;
    size_t next = 0;
    KF_CHECK(cudaGetDevice(&dev));
    env = getenv("KISSFFT_CHUNK_RAMP");
    kf_span *chunks = NULL;
    const size_t nchunks = kf_make_chunks(howmany, rows, env && env[0] == '1', &chunks);    /* opt-in: no gain measured (profiles/r02/e2e_ramp_async.txt) */
    if (!nchunks) return kf_cuda_fail(__FILE__, __LINE__, "host batch pipeline", KISS_FFT_CUDA_ENOMEM);
    const int nlanes = kf_host_lanes(nchunks, in_pinned && out_pinned);
    kf_lane lanes[KF_MAX_LANES];
    pthread_t thr[KF_MAX_LANES];
    int started[KF_MAX_LANES];
    int rc = 0;
    for (int l = 0; l < nlanes; ++l) {
        kf_lane *L = &lanes[l];
        memset(L, 0, sizeof(*L));
        started[l] = 0;
        L->fn = fn; L->cfg = cfg; L->in = (const char *)in; L->out = (char *)out;
        L->rows = rows; L->in_row_bytes = in_row_bytes; L->out_row_bytes = out_row_bytes;
        L->chunks = chunks; L->nchunks = nchunks; L->next = &next;
        L->lane = l; L->nlanes = nlanes; L->device = dev; L->in_pinned = in_pinned; L->out_pinned = out_pinned;
        if (!rc) rc = kf_ctx_acquire(&L->cx);
    }
    env = getenv("KISSFFT_HOST_ASYNC");
    if (!rc && in_pinned && out_pinned && env && env[0] == '1') {
        /* opt-in (KISSFFT_HOST_ASYNC=1; measured equal to the threaded lanes within the box-to-box spread,
         * profiles/r02/e2e_ramp_async.txt).  Pinned caller buffers: nothing for the host to do per chunk, so ONE thread enqueues every chunk up front -- chunk c on
         * the stream of slot c % nlanes (H2D, kernel, D2H in stream order; the slot's device buffers are reused in stream
         * order too) -- and waits once at the end.  Both copy engines always have queued work; the per-chunk host wake-up of
         * the threaded lanes (needed only for the bounce copies of pageable memory) is gone. */
        void *din[KF_MAX_LANES], *dout[KF_MAX_LANES];
        for (int l = 0; !rc && l < nlanes; ++l) {
            rc = kf_ctx_dev(lanes[l].cx, 0, rows * in_row_bytes, &din[l]);
            if (!rc) rc = kf_ctx_dev(lanes[l].cx, 1, rows * out_row_bytes, &dout[l]);
        }
        for (size_t c = 0; !rc && c < nchunks; ++c) {
            const int l = (int)(c % (size_t)nlanes);
            cudaStream_t st = lanes[l].cx->stream;
            const size_t first = chunks[c].first, n = chunks[c].n;
            rc = (int)cudaMemcpyAsync(din[l], (const char *)in + first * in_row_bytes, n * in_row_bytes, cudaMemcpyHostToDevice, st);
            if (!rc) rc = fn(cfg, din[l], dout[l], n, st);
            if (!rc) rc = (int)cudaMemcpyAsync((char *)out + first * out_row_bytes, dout[l], n * out_row_bytes, cudaMemcpyDeviceToHost, st);
        }
        for (int l = 0; l < nlanes; ++l) {
            const int e = (int)cudaStreamSynchronize(lanes[l].cx->stream);
            if (!rc) rc = e;
        }
    } else if (!rc) {
        for (int l = 1; l < nlanes; ++l) {
            if (pthread_create(&thr[l], NULL, kf_lane_main, &lanes[l]) == 0) started[l] = 1;
        }
        kf_lane_main(&lanes[0]);                   /* the calling thread is lane 0 */
        for (int l = 1; l < nlanes; ++l) {
            if (started[l]) pthread_join(thr[l], NULL);
            else kf_lane_main(&lanes[l]);          /* thread creation failed: run the lane here */
        }
        for (int l = 0; l < nlanes; ++l)
            if (!rc) rc = lanes[l].rc;
    }
    free(chunks);
    for (int l = 0; l < nlanes; ++l) kf_ctx_release(lanes[l].cx);
    if (rc) return kf_cuda_fail(__FILE__, __LINE__, "host batch pipeline", rc);
    return 0;
}

int kiss_fft_batch(kiss_fft_cfg cfg, const kiss_fft_cpx *in, kiss_fft_cpx *out, size_t howmany)
{
    if (!cfg || cfg->magic != KF_MAGIC_1D || !in || !out) {
        KF_ERROR("kiss_fft_batch: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    const size_t rb = sizeof(kiss_fft_cpx) * (size_t)cfg->nfft;
    return kf_host_pipeline(kf_chunk_c2c, cfg, in, out, howmany, rb, rb);
}

int kiss_fftr_batch(kiss_fftr_cfg cfg, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata, size_t howmany)
{
    if (!cfg || cfg->magic != KF_MAGIC_R || !timedata || !freqdata || cfg->substate->inverse) {
        KF_ERROR("kiss_fftr_batch: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    return kf_host_pipeline(kf_chunk_r2c, cfg, timedata, freqdata, howmany, sizeof(kiss_fft_scalar) * (size_t)cfg->nfft,
                            sizeof(kiss_fft_cpx) * ((size_t)cfg->nfft / 2 + 1));
}

int kiss_fftri_batch(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata, size_t howmany)
{
    if (!cfg || cfg->magic != KF_MAGIC_R || !timedata || !freqdata || !cfg->substate->inverse) {
        KF_ERROR("kiss_fftri_batch: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    return kf_host_pipeline(kf_chunk_c2r, cfg, freqdata, timedata, howmany, sizeof(kiss_fft_cpx) * ((size_t)cfg->nfft / 2 + 1),
                            sizeof(kiss_fft_scalar) * (size_t)cfg->nfft);
}

/* ---- the reference's transform calls -------------------------------------------------------------------- */

/* run `body` on device copies of host buffers: in (in_bytes) -> out (out_bytes) */
typedef int (*kf_dev_body)(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream);

/* bytes up to which a host buffer is bounced through the context's pinned buffer (one memcpy + an asynchronous DMA)
 * instead of being handed to cudaMemcpy as pageable memory */
#define KF_BOUNCE_MAX ((size_t)4 << 20)

static int kf_stage_through_device(kf_dev_body body, void *cfg, const void *in, size_t in_bytes, void *out, size_t out_bytes,
                                   size_t work_bytes, long long arg)
{
    kf_ctx *cx = NULL;
    KF_CHECK(kf_ctx_acquire(&cx));
    void *din = NULL, *dout = NULL, *dwork = NULL, *pin = NULL, *pout = NULL;
    cudaStream_t st = cx->stream;
    int rc = kf_ctx_dev(cx, 0, in_bytes, &din);
    if (!rc) rc = kf_ctx_dev(cx, 1, out_bytes, &dout);
    if (!rc && work_bytes) rc = kf_ctx_dev(cx, 2, work_bytes, &dwork);
    const int bounce_in = in_bytes <= KF_BOUNCE_MAX, bounce_out = out_bytes <= KF_BOUNCE_MAX;
    if (!rc && bounce_in) rc = kf_ctx_pin(cx, 0, in_bytes, &pin);
    if (!rc && bounce_out) rc = kf_ctx_pin(cx, 1, out_bytes, &pout);
    if (!rc) {
        if (bounce_in) {
            memcpy(pin, in, in_bytes);
            rc = (int)cudaMemcpyAsync(din, pin, in_bytes, cudaMemcpyHostToDevice, st);
        } else {
            rc = (int)cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, st);
        }
    }
    if (!rc) rc = body(cfg, din, dout, dwork, arg, st);
    if (!rc) rc = (int)cudaMemcpyAsync(bounce_out ? pout : out, dout, out_bytes, cudaMemcpyDeviceToHost, st);
    const int e = (int)cudaStreamSynchronize(st);      /* also after a failed body: the scratch goes back to the pool */
    if (!rc) rc = e;
    if (!rc && bounce_out) memcpy(out, pout, out_bytes);
    kf_ctx_release(cx);
    return rc;
}

static int kf_body_stride(void *cfg, const void *d_in, void *d_out, void *d_work, long long stride, void *stream)
{
    (void)d_work;
    kiss_fft_cfg c = (kiss_fft_cfg)cfg;
    return kiss_fft_batch_dev(c, (const kiss_fft_cpx *)d_in, (kiss_fft_cpx *)d_out, 1, 0, 0, (int)stride, stream);
}

void kiss_fft_stride(kiss_fft_cfg st, const kiss_fft_cpx *fin, kiss_fft_cpx *fout, int in_stride)
{
    if (!st || st->magic != KF_MAGIC_1D) {
        KF_ERROR("kiss_fft: bad cfg");
        return;
    }
    if (fout == NULL || fin == NULL) {
        KF_ERROR("fout buffer NULL."); /* kiss_fft.c:380-383 */
        return;
    }
    if (in_stride < 1) in_stride = 1;
    if (kf_is_device_ptr(fin) || kf_is_device_ptr(fout)) {
        /* device pointers: in place on the device, stream-ordered on the default stream; the kernel reads a
         * whole transform into registers/shared memory before writing it, so fin == fout is fine */
        (void)kiss_fft_batch_dev(st, fin, fout, 1, 0, 0, in_stride, NULL);
        return;
    }
    const size_t n = (size_t)st->nfft;
    const size_t in_bytes = sizeof(kiss_fft_cpx) * ((n - 1) * (size_t)in_stride + 1);
    int rc = kf_stage_through_device(kf_body_stride, st, fin, in_bytes, fout, sizeof(kiss_fft_cpx) * n, 0, in_stride);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fft_stride", rc);
}

void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout) { kiss_fft_stride(cfg, fin, fout, 1); }

static int kf_body_r2c(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream)
{
    (void)d_work; (void)arg;
    kiss_fftr_cfg c = (kiss_fftr_cfg)cfg;
    return kiss_fftr_batch_dev(c, (const kiss_fft_scalar *)d_in, (kiss_fft_cpx *)d_out, 1, 0, 0, stream);
}
static int kf_body_c2r(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream)
{
    (void)d_work; (void)arg;
    kiss_fftr_cfg c = (kiss_fftr_cfg)cfg;
    return kiss_fftri_batch_dev(c, (const kiss_fft_cpx *)d_in, (kiss_fft_scalar *)d_out, 1, 0, 0, stream);
}

void kiss_fftr(kiss_fftr_cfg st, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata)
{
    if (!st || st->magic != KF_MAGIC_R || !timedata || !freqdata) {
        KF_ERROR("kiss_fftr: bad argument");
        return;
    }
    if (st->substate->inverse) {
        KF_ERROR("kiss fft usage error: improper alloc"); /* kiss_fftr.c:69-72: logged no-op */
        return;
    }
    if (kf_is_device_ptr(timedata) || kf_is_device_ptr(freqdata)) {
        (void)kiss_fftr_batch_dev(st, timedata, freqdata, 1, 0, 0, NULL);
        return;
    }
    const size_t n = (size_t)st->nfft;
    int rc = kf_stage_through_device(kf_body_r2c, st, timedata, sizeof(kiss_fft_scalar) * n, freqdata,
                                     sizeof(kiss_fft_cpx) * (n / 2 + 1), 0, 0);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fftr", rc);
}

void kiss_fftri(kiss_fftr_cfg st, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata)
{
    if (!st || st->magic != KF_MAGIC_R || !timedata || !freqdata) {
        KF_ERROR("kiss_fftri: bad argument");
        return;
    }
    if (st->substate->inverse == 0) {
        KF_ERROR("kiss fft usage error: improper alloc"); /* kiss_fftr.c:124-127 */
        return;
    }
    if (kf_is_device_ptr(timedata) || kf_is_device_ptr(freqdata)) {
        (void)kiss_fftri_batch_dev(st, freqdata, timedata, 1, 0, 0, NULL);
        return;
    }
    const size_t n = (size_t)st->nfft;
    int rc = kf_stage_through_device(kf_body_c2r, st, freqdata, sizeof(kiss_fft_cpx) * (n / 2 + 1), timedata,
                                     sizeof(kiss_fft_scalar) * n, 0, 0);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fftri", rc);
}

static int kf_body_nd(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream)
{
    (void)arg;
    kiss_fftnd_cfg st = (kiss_fftnd_cfg)cfg;
    if (!d_work) return kf_axes_inlayout(st, st->ndims, 1, (const kiss_fft_cpx *)d_in, (kiss_fft_cpx *)d_out, stream);
#ifndef FIXED_POINT
    if (kf_fftnd3_permuted_wanted(st, (const kiss_fft_cpx *)d_in, (const kiss_fft_cpx *)d_out))
        return kf_fftnd3_permuted(st, (const kiss_fft_cpx *)d_in, (kiss_fft_cpx *)d_out, (kiss_fft_cpx *)d_work, stream);
#endif
    return kf_fftnd_dev_locked(st, (const kiss_fft_cpx *)d_in, (kiss_fft_cpx *)d_out, (kiss_fft_cpx *)d_work, stream);
}
static int kf_body_ndr(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream)
{
    (void)d_work; (void)arg;
    return kiss_fftndr_dev((kiss_fftndr_cfg)cfg, (const kiss_fft_scalar *)d_in, (kiss_fft_cpx *)d_out, stream);
}
static int kf_body_ndri(void *cfg, const void *d_in, void *d_out, void *d_work, long long arg, void *stream)
{
    (void)d_work; (void)arg;
    return kiss_fftndri_dev((kiss_fftndr_cfg)cfg, (const kiss_fft_cpx *)d_in, (kiss_fft_scalar *)d_out, stream);
}

void kiss_fftnd(kiss_fftnd_cfg st, const kiss_fft_cpx *fin, kiss_fft_cpx *fout)
{
    if (!st || st->magic != KF_MAGIC_ND || !fin || !fout) {
        KF_ERROR("kiss_fftnd: bad argument");
        return;
    }
    if (kf_is_device_ptr(fin) || kf_is_device_ptr(fout)) {
        (void)kiss_fftnd_dev(st, fin, fout, NULL, NULL);
        return;
    }
    const size_t bytes = sizeof(kiss_fft_cpx) * (size_t)st->dimprod;
    int inlay = st->ndims >= 2 && kf_axes_inlayout_ok(st, st->ndims, 1);      /* then no work buffer is needed */
#ifndef FIXED_POINT
    if (kf_fftnd3_permuted_wanted(st, fin, NULL)) inlay = 0;
#endif
    int rc = kf_stage_through_device(kf_body_nd, st, fin, bytes, fout, bytes, inlay ? 0 : bytes, 0);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fftnd", rc);
}

void kiss_fftndr(kiss_fftndr_cfg st, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata)
{
    if (!st || st->magic != KF_MAGIC_NDR || !timedata || !freqdata) {
        KF_ERROR("kiss_fftndr: bad argument");
        return;
    }
    if (kf_is_device_ptr(timedata) || kf_is_device_ptr(freqdata)) {
        (void)kiss_fftndr_dev(st, timedata, freqdata, NULL);
        return;
    }
    const size_t nrbins = (size_t)st->dimReal / 2 + 1;
    const size_t in_bytes = sizeof(kiss_fft_scalar) * (size_t)st->dimOther * (size_t)st->dimReal;
    const size_t out_bytes = sizeof(kiss_fft_cpx) * (size_t)st->dimOther * nrbins;
    int rc = kf_stage_through_device(kf_body_ndr, st, timedata, in_bytes, freqdata, out_bytes, 0, 0);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fftndr", rc);
}

void kiss_fftndri(kiss_fftndr_cfg st, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata)
{
    if (!st || st->magic != KF_MAGIC_NDR || !timedata || !freqdata) {
        KF_ERROR("kiss_fftndri: bad argument");
        return;
    }
    if (kf_is_device_ptr(timedata) || kf_is_device_ptr(freqdata)) {
        (void)kiss_fftndri_dev(st, freqdata, timedata, NULL);
        return;
    }
    const size_t nrbins = (size_t)st->dimReal / 2 + 1;
    const size_t out_bytes = sizeof(kiss_fft_scalar) * (size_t)st->dimOther * (size_t)st->dimReal;
    const size_t in_bytes = sizeof(kiss_fft_cpx) * (size_t)st->dimOther * nrbins;
    int rc = kf_stage_through_device(kf_body_ndri, st, freqdata, in_bytes, timedata, out_bytes, 0, 0);
    if (rc) kf_cuda_fail(__FILE__, __LINE__, "kiss_fftndri", rc);
}

void kiss_fft_cuda_force_generic(int on) { kfcu_force_generic(on); }
void kiss_fft_cuda_set_grid_limit(int max_ctas) { kfcu_set_grid_limit(max_ctas); }

int kiss_fft_cuda_plan_kind(int nfft)
{
    if (nfft <= 0) return -1;
    if (kfcu_has_fused(nfft, KFCU_C2C)) return 1;
    return nfft <= kfcu_generic_max_nfft() ? 0 : 2;
}

/* ---- kfc: cfg cache (reference kfc.c:13-83) --------------------------------------------------------------- */
typedef struct kfc_entry {
    int nfft, inverse;
    kiss_fft_cfg cfg;
    struct kfc_entry *next;
} kfc_entry;
static kfc_entry *g_kfc = NULL;
static pthread_mutex_t g_kfc_lock = PTHREAD_MUTEX_INITIALIZER;

static kiss_fft_cfg kfc_find(int nfft, int inverse)
{
    pthread_mutex_lock(&g_kfc_lock);
    kfc_entry *e;
    for (e = g_kfc; e; e = e->next)
        if (e->nfft == nfft && e->inverse == inverse) break;
    if (!e) {
        e = (kfc_entry *)malloc(sizeof(*e));
        if (e) {
            e->nfft = nfft;
            e->inverse = inverse;
            e->cfg = kiss_fft_alloc(nfft, inverse, NULL, NULL);
            if (!e->cfg) {
                free(e);
                e = NULL;
            } else {
                e->next = g_kfc;
                g_kfc = e;
            }
        }
    }
    pthread_mutex_unlock(&g_kfc_lock);
    return e ? e->cfg : NULL;
}

void kfc_fft(int nfft, const kiss_fft_cpx *fin, kiss_fft_cpx *fout)
{
    kiss_fft_cfg cfg = kfc_find(nfft, 0);
    if (cfg) kiss_fft(cfg, fin, fout);
}

void kfc_ifft(int nfft, const kiss_fft_cpx *fin, kiss_fft_cpx *fout)
{
    kiss_fft_cfg cfg = kfc_find(nfft, 1);
    if (cfg) kiss_fft(cfg, fin, fout);
}

void kfc_cleanup(void)
{
    pthread_mutex_lock(&g_kfc_lock);
    for (kfc_entry *e = g_kfc; e;) {
        kfc_entry *n = e->next;
        kiss_fft_free(e->cfg);
        free(e);
        e = n;
    }
    g_kfc = NULL;
    pthread_mutex_unlock(&g_kfc_lock);
}

/* ---- fast convolution (reference tools/kiss_fastfir.c: complex-sample build and REAL_FASTFIR build) ------------------ */
#ifndef FIXED_POINT
#define KF_MAGIC_FC 0x4b464643u
#define KF_MAGIC_FCR 0x4b464352u
struct kiss_fastconv_state {
    uint32_t magic;
    int nfft, ngood, nbins; /* nbins = nfft (complex samples) or nfft/2+1 (real samples, kiss_fastfir.c:91-95) */
    kiss_fft_cfg fwd, inv;
    kiss_fftr_cfg rfwd, rinv;
    void *d_resp;           /* scaled frequency response on the device (current device at alloc time) */
    int device;
};

/* the reference's default size: next power of two at least twice the impulse response (kiss_fastfir.c:76-84) */
static size_t kf_fastconv_default_nfft(size_t n_imp_resp)
{
    size_t i = n_imp_resp - 1, nfft = 2;
    do {
        nfft <<= 1;
    } while (i >>= 1);
    return nfft;
}

kiss_fastconv_cfg kiss_fastconv_alloc(const kiss_fft_cpx *imp_resp, size_t n_imp_resp, size_t *pnfft)
{
    if (!imp_resp || n_imp_resp < 1) return NULL;
    size_t nfft = pnfft ? *pnfft : 0;
    if (nfft == 0) nfft = kf_fastconv_default_nfft(n_imp_resp);
    if (n_imp_resp > nfft || nfft > (size_t)INT32_MAX) return NULL;
    if (pnfft) *pnfft = nfft;
    kiss_fastconv_cfg st = (kiss_fastconv_cfg)calloc(1, sizeof(*st));
    kiss_fft_cpx *tmp = (kiss_fft_cpx *)calloc(nfft, sizeof(kiss_fft_cpx));
    kiss_fft_cpx *resp = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * nfft);
    if (!st || !tmp || !resp) { free(st); free(tmp); free(resp); return NULL; }
    st->magic = KF_MAGIC_FC;
    st->nfft = (int)nfft;
    st->nbins = (int)nfft;
    st->ngood = (int)(nfft - n_imp_resp + 1);
    st->fwd = kiss_fft_alloc((int)nfft, 0, NULL, NULL);
    st->inv = kiss_fft_alloc((int)nfft, 1, NULL, NULL);
    int ok = st->fwd && st->inv;
    if (ok) {
        /* zero pad in the middle to left-rotate the impulse response: the scrap samples end up at the END of each
         * inverse-transformed block (kiss_fastfir.c:139-147) */
        tmp[0] = imp_resp[n_imp_resp - 1];
        for (size_t i = 0; i + 1 < n_imp_resp; ++i) tmp[nfft - n_imp_resp + 1 + i] = imp_resp[i];
        kiss_fft(st->fwd, tmp, resp); /* on the GPU, through the library itself */
        const float scale = 1.0f / (float)nfft; /* kiss_fastfir.c:152-162 */
        for (size_t i = 0; i < nfft; ++i) {
            resp[i].r *= scale;
            resp[i].i *= scale;
        }
        ok = cudaGetDevice(&st->device) == cudaSuccess && cudaMalloc(&st->d_resp, sizeof(kiss_fft_cpx) * nfft) == cudaSuccess &&
             cudaMemcpy(st->d_resp, resp, sizeof(kiss_fft_cpx) * nfft, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    free(tmp);
    free(resp);
    if (!ok) {
        kiss_fastconv_free(st);
        return NULL;
    }
    return st;
}

/* REAL_FASTFIR (kiss_fastfir.c:26-33, 91-95): real samples and impulse response, kiss_fftr / kiss_fftri, nfft/2+1 bins */
kiss_fastconv_cfg kiss_fastconvr_alloc(const kiss_fft_scalar *imp_resp, size_t n_imp_resp, size_t *pnfft)
{
    if (!imp_resp || n_imp_resp < 1) return NULL;
    size_t nfft = pnfft ? *pnfft : 0;
    if (nfft == 0) nfft = kf_fastconv_default_nfft(n_imp_resp);
    if (n_imp_resp > nfft || (nfft & 1) || nfft > (size_t)INT32_MAX) return NULL;
    if (pnfft) *pnfft = nfft;
    const size_t nbins = nfft / 2 + 1;
    kiss_fastconv_cfg st = (kiss_fastconv_cfg)calloc(1, sizeof(*st));
    kiss_fft_scalar *tmp = (kiss_fft_scalar *)calloc(nfft, sizeof(kiss_fft_scalar));
    kiss_fft_cpx *resp = (kiss_fft_cpx *)malloc(sizeof(kiss_fft_cpx) * nbins);
    if (!st || !tmp || !resp) { free(st); free(tmp); free(resp); return NULL; }
    st->magic = KF_MAGIC_FCR;
    st->nfft = (int)nfft;
    st->nbins = (int)nbins;
    st->ngood = (int)(nfft - n_imp_resp + 1);
    st->rfwd = kiss_fftr_alloc((int)nfft, 0, NULL, NULL);
    st->rinv = kiss_fftr_alloc((int)nfft, 1, NULL, NULL);
    int ok = st->rfwd && st->rinv;
    if (ok) {
        tmp[0] = imp_resp[n_imp_resp - 1];
        for (size_t i = 0; i + 1 < n_imp_resp; ++i) tmp[nfft - n_imp_resp + 1 + i] = imp_resp[i];
        kiss_fftr(st->rfwd, tmp, resp);
        const float scale = 1.0f / (float)nfft;
        for (size_t i = 0; i < nbins; ++i) {
            resp[i].r *= scale;
            resp[i].i *= scale;
        }
        ok = cudaGetDevice(&st->device) == cudaSuccess && cudaMalloc(&st->d_resp, sizeof(kiss_fft_cpx) * nbins) == cudaSuccess &&
             cudaMemcpy(st->d_resp, resp, sizeof(kiss_fft_cpx) * nbins, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    free(tmp);
    free(resp);
    if (!ok) {
        kiss_fastconv_free(st);
        return NULL;
    }
    return st;
}

void kiss_fastconv_free(kiss_fastconv_cfg st)
{
    if (!st) return;
    if (st->d_resp) cudaFree(st->d_resp);
    kiss_fft_free(st->fwd);
    kiss_fft_free(st->inv);
    kiss_fftr_free(st->rfwd);
    kiss_fftr_free(st->rinv);
    free(st);
}

size_t kiss_fastconv_block_advance(kiss_fastconv_cfg st) { return st ? (size_t)st->ngood : 0; }
size_t kiss_fastconv_nfft(kiss_fastconv_cfg st) { return st ? (size_t)st->nfft : 0; }

/* fastconv1buf composed from separate launches, for lengths without a fused kernel and for the real-sample build:
 * dense copies of the overlapping blocks -> batched forward transform -> pointwise product with the response -> batched
 * inverse transform -> the ngood valid samples of every block.  Any nfft the transforms accept (four-step beyond 4096 x
 * 4); scratch comes from a staging context, so this path waits for the stream before it returns. */
static int kf_fastconv_unfused(kiss_fastconv_cfg st, const void *d_in, void *d_out, size_t nblocks, void *stream)
{
    const int real = st->magic == KF_MAGIC_FCR;
    const size_t nfft = (size_t)st->nfft, nbins = (size_t)st->nbins, ngood = (size_t)st->ngood;
    const size_t ssz = real ? sizeof(kiss_fft_scalar) : sizeof(kiss_fft_cpx);
    kf_ctx *cx = NULL;
    KF_CHECK(kf_ctx_acquire(&cx));
    void *blk = NULL, *spec = NULL;
    int rc = kf_ctx_dev(cx, 0, ssz * nfft * nblocks, &blk);
    if (!rc) rc = kf_ctx_dev(cx, 1, sizeof(kiss_fft_cpx) * nbins * nblocks, &spec);
    if (!rc) rc = kfcu_gather_blocks(d_in, blk, (long long)nblocks, (int)nfft, (long long)ngood, real, stream);
    if (!rc) {
        if (real) rc = kiss_fftr_batch_dev(st->rfwd, (const kiss_fft_scalar *)blk, (kiss_fft_cpx *)spec, nblocks, nfft, nbins, stream);
        else rc = kiss_fft_batch_dev(st->fwd, (const kiss_fft_cpx *)blk, (kiss_fft_cpx *)spec, nblocks, nfft, nfft, 1, stream);
    }
    if (!rc) rc = kfcu_cmul_rows(spec, st->d_resp, (long long)nblocks, (int)nbins, stream);
    if (!rc) {
        if (real) rc = kiss_fftri_batch_dev(st->rinv, (const kiss_fft_cpx *)spec, (kiss_fft_scalar *)blk, nblocks, nbins, nfft, stream);
        else rc = kiss_fft_batch_dev(st->inv, (const kiss_fft_cpx *)spec, (kiss_fft_cpx *)blk, nblocks, nfft, nfft, 1, stream);
    }
    if (!rc)
        rc = (int)cudaMemcpy2DAsync(d_out, ssz * ngood, blk, ssz * nfft, ssz * ngood, nblocks, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    const int e = (int)cudaStreamSynchronize((cudaStream_t)stream);
    kf_ctx_release(cx);
    return rc ? rc : e;
}

/* kff_nocopy (kiss_fastfir.c:191-206) on device buffers: all complete blocks of the n input samples. */
int kiss_fastconv_dev(kiss_fastconv_cfg st, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t n, size_t *nprocessed,
                      void *stream)
{
    if (!st || st->magic != KF_MAGIC_FC || !d_in || !d_out) {
        KF_ERROR("kiss_fastconv_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    size_t nblocks = 0;
    if (n >= (size_t)st->nfft) nblocks = (n - (size_t)st->nfft) / (size_t)st->ngood + 1;
    if (nprocessed) *nprocessed = nblocks * (size_t)st->ngood;
    if (nblocks == 0) return 0;
    if (kfcu_has_fastconv(st->nfft)) {
        const kf_devplan *pf, *pi;
        KF_CHECK(kf_get_devplan(st->fwd, NULL, &pf));
        KF_CHECK(kf_get_devplan(st->inv, NULL, &pi));
        KF_CHECK(kfcu_fastconv((kfcu_plan *)&pf->plan, (kfcu_plan *)&pi->plan, d_in, d_out, (long long)nblocks, st->ngood, st->d_resp,
                               stream));
        return 0;
    }
    KF_CHECK(kf_fastconv_unfused(st, d_in, d_out, nblocks, stream));
    return 0;
}

int kiss_fastconvr_dev(kiss_fastconv_cfg st, const kiss_fft_scalar *d_in, kiss_fft_scalar *d_out, size_t n, size_t *nprocessed,
                       void *stream)
{
    if (!st || st->magic != KF_MAGIC_FCR || !d_in || !d_out) {
        KF_ERROR("kiss_fastconvr_dev: bad argument");
        return KISS_FFT_CUDA_EINVAL;
    }
    size_t nblocks = 0;
    if (n >= (size_t)st->nfft) nblocks = (n - (size_t)st->nfft) / (size_t)st->ngood + 1;
    if (nprocessed) *nprocessed = nblocks * (size_t)st->ngood;
    if (nblocks == 0) return 0;
    KF_CHECK(kf_fastconv_unfused(st, d_in, d_out, nblocks, stream));
    return 0;
}
#endif
