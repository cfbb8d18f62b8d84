/*
 * kf_mgpu.c -- kiss_fftnd over several GPUs, one process per GPU, behind a C-ABI (include/kiss_fft_cuda.h:
 * kiss_fftnd_mgpu_get_id / _alloc / _exec / _free).
 *
 * The reference transforms ONE array with kiss_fftnd (kiss_fftnd.c:156-188) and has no distributed code.  Here the
 * d0 x d1 x d2 array is split into slabs of P = d0/G planes per rank and needs a single exchange:
 *
 *   A. rows (axis 2) of the local slab, in place                      kiss_fft_batch_dev          [P][d1][d2]
 *   B. columns (axis 1) of every local plane, written transposed and already sorted by destination rank -- the k2
 *      columns are cut into `nchunks` chunks of cw = (d2/G)/nchunks so that the exchange and step C pipeline:
 *      chunk j of destination s holds k2 in s*d2/G + [j*cw, (j+1)*cw)   kiss_fft_planes_pass_peers_dev
 *   X. all-to-all: block (s, j) goes to rank s.  Two implementations:
 *        - NCCL: step B writes a local send buffer, grouped ncclSend/ncclRecv per chunk on a communication stream
 *          (libnccl is dlopen()ed; the communicator is built from the id rank 0 hands out);
 *        - peer stores (KISS_FFT_MGPU_P2P): the receive buffers of all ranks are mapped into every process with CUDA
 *          IPC and step B's kernel stores its rows straight into them over NVLink -- no send buffer, no collective;
 *          a flag per (chunk, source rank) in the receiver's memory says "landed".
 *   C. axis 0 of chunk j, as soon as chunk j has arrived from every rank  kiss_fft_axis_pass_dev  [d2/G][d1][d0]
 *
 * Pipeline: the k2 columns are cut into chunks (2; the last one narrower, kf_set_chunks) and the local planes into groups (2).
 * A(0) runs on the caller's stream; B(i, 0) follows A(i), so the link-bound stores of the first chunk overlap the rows of the
 * next plane group; then B(., 1), ... while C(j-1) runs as soon as chunk j-1 has landed from every rank.  Exposed: A(0), the
 * exchange itself, and C of the last chunk.  Overlap between these persistent kernels needs the SMs partitioned by hand
 * (kf_make_partition: CUDA green contexts) -- B in one partition, A(i > 0) and C(j < last) in the other, A(0) and the last C
 * on the whole device; without green contexts the launches of B are capped instead (KISSFFT_MGPU_B_CTAS).
 * kiss_fftnd_mgpu_tune / _knob / _trace are the tuning aids the sweeps in profiles/r02/mgpu_* were made with.
 *
 * Receive buffer layout.  NCCL: [chunk j][source rank r][P][cw][d1] -- one contiguous piece per (source, chunk), and chunk
 * j is a dense (G*P) x (cw*d1) matrix for step C (kiss_fft_axis_pass_dev).  Peer stores: [chunk j][c][i0 = r*P + p][d1] --
 * the source's kernel places every row itself, so all d0 planes of one k2 column sit together and step C is a
 * plane-batched pass (kiss_fft_planes_pass_dev, cw planes of d0 x d1) whose rows are 8 KiB apart instead of cw*d1*8
 * bytes: address translation stays within a few pages per tile and the tensor-map input ring applies
 * (profiles/r02/tune_r2c_f32_col1024_*: 0.88 of the HBM peak against 0.6 for rows on different pages).  Output: X[k0][k1][k2] stored as out[k2 - r*d2/G][k1][k0]
 * on rank r ("transposed out", distributed along k2), the usual contract of slab FFTs.
 *
 * Host code is C; the only CUDA code it needs beyond the library's own entry points are the two flag kernels
 * (kf_launch.cu: kfcu_peer_signal / kfcu_peer_wait).
 */
#include <cuda.h>               /* types of the green-context API only; entry points come from cudaGetDriverEntryPoint */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/kiss_fft_cuda.h"
#include "kf_internal.h"

/* ---- the few NCCL declarations used (nccl.h: 2.19+ ABI); the library is loaded at run time --------------------- */
typedef struct { char internal[128]; } kf_nccl_id;
typedef void *kf_nccl_comm;
enum { KF_NCCL_UINT8 = 1, KF_NCCL_SUCCESS = 0 };
typedef struct {
    void *lib;
    int (*GetUniqueId)(kf_nccl_id *);
    int (*CommInitRank)(kf_nccl_comm *, int, kf_nccl_id, int);
    int (*CommDestroy)(kf_nccl_comm);
    int (*Send)(const void *, size_t, int, int, kf_nccl_comm, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, kf_nccl_comm, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, kf_nccl_comm, cudaStream_t);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char *(*GetErrorString)(int);
} kf_nccl_api;

static kf_nccl_api g_nccl;

static int kf_nccl_load(void)
{
    if (g_nccl.lib) return 0;
    /* a copy already in the process (e.g. the one PyTorch bundles) is preferred over the system one */
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    kf_nccl_api a;
    memset(&a, 0, sizeof(a));
    a.lib = h;
    *(void **)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void **)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void **)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void **)&a.Send = dlsym(h, "ncclSend");
    *(void **)&a.Recv = dlsym(h, "ncclRecv");
    *(void **)&a.AllGather = dlsym(h, "ncclAllGather");
    *(void **)&a.GroupStart = dlsym(h, "ncclGroupStart");
    *(void **)&a.GroupEnd = dlsym(h, "ncclGroupEnd");
    *(void **)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Send || !a.Recv || !a.AllGather || !a.GroupStart || !a.GroupEnd)
        return -1;
    g_nccl = a;
    return 0;
}


/* ---- SM partitions (CUDA green contexts, driver API >= 12.4) ------------------------------------------------------
 * Two persistent kernels on two streams do not overlap by themselves: whichever starts first fills every SM, and a
 * capped grid does not help either, because the CTA scheduler spreads the capped kernel's small CTAs over all SMs and the
 * other kernel's 198 KB CTAs then fit nowhere (timeline: profiles/r02/mgpu_g2_timeline_before_partition.txt).  A green
 * context owns a fixed set of SMs; a stream created in it launches only there.  The link-bound launches (B: column pass
 * whose stores cross NVLink) get one partition, the HBM-bound launches that run beside them (A(i > 0), C(j < last)) the
 * other; A(0) and the last C run in the caller's context on the whole device.  Memory, events and kernels are shared
 * with the primary context, so nothing else changes. */
typedef struct {
    int tried, ok;
    CUresult (*DeviceGet)(CUdevice *, int);
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType);
    CUresult (*DevSmResourceSplitByCount)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    CUresult (*GreenCtxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    CUresult (*GreenCtxDestroy)(CUgreenCtx);
    CUresult (*GreenCtxStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int);
} kf_drv_api;
static kf_drv_api g_drv;

static int kf_drv_load(void)
{
    if (g_drv.tried) return g_drv.ok ? 0 : -1;
    kf_drv_api a;
    memset(&a, 0, sizeof(a));
    a.tried = 1;
    struct { const char *name; void **fn; } want[] = {
        {"cuDeviceGet", (void **)&a.DeviceGet},
        {"cuDeviceGetDevResource", (void **)&a.DeviceGetDevResource},
        {"cuDevSmResourceSplitByCount", (void **)&a.DevSmResourceSplitByCount},
        {"cuDevResourceGenerateDesc", (void **)&a.DevResourceGenerateDesc},
        {"cuGreenCtxCreate", (void **)&a.GreenCtxCreate},
        {"cuGreenCtxDestroy", (void **)&a.GreenCtxDestroy},
        {"cuGreenCtxStreamCreate", (void **)&a.GreenCtxStreamCreate},
    };
    a.ok = 1;
    for (size_t i = 0; i < sizeof(want) / sizeof(want[0]); ++i) {
        enum cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint(want[i].name, want[i].fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*want[i].fn) {
            cudaGetLastError();
            a.ok = 0;
            break;
        }
    }
    g_drv = a;
    return a.ok ? 0 : -1;
}

#define KF_MAGIC_MGPU 0x4b464d47u
#define KF_MGPU_MAXRANKS 16
#define KF_MGPU_MAXCHUNKS 8
#define KF_MGPU_MAXTRACE 96
#define KF_FLAG_BYTES 4096 /* (MAXCHUNKS + 1) slots x MAXRANKS unsigned, rounded up */

struct kiss_fftnd_mgpu_state {
    uint32_t magic;
    int d0, d1, d2, inverse, rank, nranks, device;
    int planes, cols, nchunks, cw; /* P, C = d2/G, chunks of the k2 columns, columns per chunk (uniform split) */
    int coff[KF_MGPU_MAXCHUNKS + 1]; /* first column of every chunk; the last chunk may be narrower (tail16) so that the exposed last C is short */
    int tail16;                    /* share of the last chunk in sixteenths of C (0 = equal chunks) */
    int ndims;                     /* 3, or 2 (then d1 == 1) */
    int b_prio;                    /* s_b is a highest-priority stream */
    int trace, ntr;                /* tuning aid: timed events around every launch of the last exec (kiss_fftnd_mgpu_trace) */
    cudaEvent_t tr_ev[KF_MGPU_MAXTRACE][2], tr_t0;
    char tr_name[KF_MGPU_MAXTRACE][12];
    int gc_on, link_sms, rest_sms, dev_sms;  /* green-context SM partition: SMs of the link-bound / of the HBM-bound partition */
    CUgreenCtx gc_link, gc_rest;
    cudaStream_t gs_b, gs_a, gs_c; /* B in the link partition; A(i > 0) and C(j < last) in the other */
    cudaEvent_t ev_adone, ev_done2;
    int ac_reserve;                /* SMs the HBM-bound launches (A, C) leave to the concurrent link-bound / NCCL kernels */
    int pchunks, b_ctas;           /* groups of local planes (pipelining of A against B); CTA cap of link-bound B launches */
    unsigned flags;
    int p2p;                       /* peer-store exchange active */
    kiss_fft_cfg cfg0, cfg1, cfg2;
    kf_nccl_comm comm;
    char *recv_base;               /* one allocation: receive buffer followed by the flag words */
    size_t recv_bytes;
    kiss_fft_cpx *send;            /* NCCL path only */
    kiss_fft_cpx *work;            /* reference-order mode only: the slab after the axis-0 pass, [d1][C][d0] */
    char *peer_base[KF_MGPU_MAXRANKS]; /* every rank's recv_base as mapped here (own entry = recv_base) */
    unsigned **d_peer_flags;       /* device copy of the G flag-array pointers */
    unsigned epoch;
    cudaStream_t s_comm, s_c, s_b;
    cudaEvent_t ev_start, ev_a[KF_MGPU_MAXCHUNKS], ev_b[KF_MGPU_MAXCHUNKS], ev_x[KF_MGPU_MAXCHUNKS], ev_done, ev_bdone;
    char err[256];
};

static __thread char tls_mgpu_err[256];
const char *kiss_fftnd_mgpu_last_error(void) { return tls_mgpu_err; }

static int kf_fail(kiss_fftnd_mgpu_cfg st, const char *what, int code, int is_nccl)
{
    const char *msg = is_nccl ? (g_nccl.GetErrorString ? g_nccl.GetErrorString(code) : "nccl error")
                              : (code > 0 ? cudaGetErrorString((cudaError_t)code) : kiss_fft_cuda_last_error());
    snprintf(tls_mgpu_err, sizeof(tls_mgpu_err), "%s: %s %d (%s)", what, is_nccl ? "NCCL" : "error", code, msg ? msg : "");
    if (st) memcpy(st->err, tls_mgpu_err, sizeof(st->err));
#ifndef NDEBUG
    fprintf(stderr, "[ERROR] %s:%d %s\n", __FILE__, __LINE__, tls_mgpu_err);
#endif
    return code ? code : KISS_FFT_CUDA_EINVAL;
}
#define CU(expr)                                                      \
    do {                                                              \
        int rc_ = (int)(expr);                                        \
        if (rc_ != 0) return kf_fail(st, #expr, rc_, 0);              \
    } while (0)
#define NC(expr)                                                      \
    do {                                                              \
        int rc_ = (int)(expr);                                        \
        if (rc_ != KF_NCCL_SUCCESS) return kf_fail(st, #expr, rc_, 1); \
    } while (0)

int kiss_fftnd_mgpu_get_id(void *id)
{
    if (!id) return KISS_FFT_CUDA_EINVAL;
    if (kf_nccl_load() != 0) return kf_fail(NULL, "libnccl.so.2 could not be loaded", KISS_FFT_CUDA_EINVAL, 0);
    kf_nccl_id nid;
    const int rc = g_nccl.GetUniqueId(&nid);
    if (rc != KF_NCCL_SUCCESS) return kf_fail(NULL, "ncclGetUniqueId", rc, 1);
    memcpy(id, &nid, KISS_FFT_MGPU_ID_BYTES);
    return 0;
}

static unsigned *flag_ptr(const struct kiss_fftnd_mgpu_state *st, int r) { return (unsigned *)(st->peer_base[r] + st->recv_bytes); }

/* map every rank's receive buffer into this process (CUDA IPC); handles travel through an NCCL all-gather */
static int kf_mgpu_map_peers(kiss_fftnd_mgpu_cfg st)
{
    cudaIpcMemHandle_t mine, *all = NULL;
    void *d_h = NULL;
    const int G = st->nranks;
    int rc = (int)cudaIpcGetMemHandle(&mine, st->recv_base);
    if (rc) { cudaGetLastError(); return rc; }
    all = (cudaIpcMemHandle_t *)malloc(sizeof(*all) * (size_t)G);
    if (!all) return KISS_FFT_CUDA_ENOMEM;
    rc = (int)cudaMalloc(&d_h, sizeof(mine) * (size_t)(G + 1));
    if (!rc) rc = (int)cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    if (!rc) {
        const int n = g_nccl.AllGather(d_h, (char *)d_h + sizeof(mine), sizeof(mine), KF_NCCL_UINT8, st->comm, (cudaStream_t)0);
        if (n != KF_NCCL_SUCCESS) rc = KISS_FFT_CUDA_EINVAL;
    }
    if (!rc) rc = (int)cudaStreamSynchronize((cudaStream_t)0);
    if (!rc) rc = (int)cudaMemcpy(all, (char *)d_h + sizeof(mine), sizeof(mine) * (size_t)G, cudaMemcpyDeviceToHost);
    for (int r = 0; !rc && r < G; ++r) {
        if (r == st->rank) {
            st->peer_base[r] = st->recv_base;
        } else {
            void *p = NULL;
            rc = (int)cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
            if (rc) cudaGetLastError();
            st->peer_base[r] = (char *)p;
        }
    }
    if (d_h) cudaFree(d_h);
    free(all);
    if (rc) {
        for (int r = 0; r < G; ++r) {
            if (r != st->rank && st->peer_base[r]) cudaIpcCloseMemHandle(st->peer_base[r]);
            st->peer_base[r] = NULL;
        }
        return rc;
    }
    unsigned *ptrs[KF_MGPU_MAXRANKS];
    for (int r = 0; r < G; ++r) ptrs[r] = flag_ptr(st, r);
    rc = (int)cudaMalloc((void **)&st->d_peer_flags, sizeof(unsigned *) * (size_t)G);
    if (!rc) rc = (int)cudaMemcpy(st->d_peer_flags, ptrs, sizeof(unsigned *) * (size_t)G, cudaMemcpyHostToDevice);
    return rc;
}

/* The stream of the link-bound B launches gets the highest priority: when a kernel of another stream drains, B's few
 * CTAs are placed first and the HBM-bound kernel fills the rest of the device, so the NVLink stores start as early as
 * their inputs allow instead of queueing behind the next full-grid launch. */
static int kf_make_b_stream(kiss_fftnd_mgpu_cfg st, int prio)
{
    int lo = 0, hi = 0;
    if (st->s_b) { cudaStreamDestroy(st->s_b); st->s_b = NULL; }
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { cudaGetLastError(); prio = 0; }
    st->b_prio = prio;
    return (int)cudaStreamCreateWithPriority(&st->s_b, cudaStreamNonBlocking, prio ? hi : lo);
}

static void kf_drop_partition(kiss_fftnd_mgpu_cfg st)
{
    if (st->gs_b) cudaStreamDestroy(st->gs_b);
    if (st->gs_a) cudaStreamDestroy(st->gs_a);
    if (st->gs_c) cudaStreamDestroy(st->gs_c);
    st->gs_b = st->gs_a = st->gs_c = NULL;
    if (st->gc_link) g_drv.GreenCtxDestroy(st->gc_link);
    if (st->gc_rest) g_drv.GreenCtxDestroy(st->gc_rest);
    st->gc_link = st->gc_rest = NULL;
    st->gc_on = st->link_sms = st->rest_sms = 0;
}

/* split the device into `want_link` SMs (rounded up to the hardware's granularity, 8 on sm_90+) and the rest; 0 = no
 * partition.  Failure of any step leaves the cfg on the unpartitioned pipeline (not an error). */
static int kf_make_partition(kiss_fftnd_mgpu_cfg st, int want_link)
{
    if (st->gc_on || st->gs_b) kf_drop_partition(st);
    if (want_link <= 0 || kf_drv_load() != 0) return -1;
    CUdevice dev;
    CUdevResource all, grp, rest;
    CUdevResourceDesc d_link = NULL, d_rest = NULL;
    unsigned int n = 1;
    memset(&all, 0, sizeof(all)); memset(&grp, 0, sizeof(grp)); memset(&rest, 0, sizeof(rest));
    int ok = g_drv.DeviceGet(&dev, st->device) == CUDA_SUCCESS &&
             g_drv.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS;
    if (ok && (unsigned)want_link + 8 > all.sm.smCount) ok = 0;
    ok = ok && g_drv.DevSmResourceSplitByCount(&grp, &n, &all, &rest, 0, (unsigned)want_link) == CUDA_SUCCESS && n == 1 &&
         grp.type == CU_DEV_RESOURCE_TYPE_SM && rest.type == CU_DEV_RESOURCE_TYPE_SM && rest.sm.smCount >= 8;
    ok = ok && g_drv.DevResourceGenerateDesc(&d_link, &grp, 1) == CUDA_SUCCESS && g_drv.DevResourceGenerateDesc(&d_rest, &rest, 1) == CUDA_SUCCESS;
    ok = ok && g_drv.GreenCtxCreate(&st->gc_link, d_link, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS &&
         g_drv.GreenCtxCreate(&st->gc_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS;
    ok = ok && g_drv.GreenCtxStreamCreate((CUstream *)&st->gs_b, st->gc_link, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
         g_drv.GreenCtxStreamCreate((CUstream *)&st->gs_a, st->gc_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS &&
         g_drv.GreenCtxStreamCreate((CUstream *)&st->gs_c, st->gc_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
    if (!ok) { cudaGetLastError(); kf_drop_partition(st); return -1; }
    st->dev_sms = (int)all.sm.smCount;
    st->link_sms = (int)grp.sm.smCount;
    st->rest_sms = (int)rest.sm.smCount;
    st->gc_on = 1;
    return 0;
}

static void kf_set_chunks(kiss_fftnd_mgpu_cfg st, int want)
{
    if (want > KF_MGPU_MAXCHUNKS) want = KF_MGPU_MAXCHUNKS;
    if (st->nranks == 1 || want < 1) want = 1;
    st->nchunks = 1;
    for (int c = want; c >= 1; --c)
        if (st->cols % c == 0 && ((st->cols / c) % 16 == 0 || c == 1)) { st->nchunks = c; break; }
    if (st->flags & KISS_FFT_MGPU_REFERENCE_ORDER) st->nchunks = 1;     /* one exchange block per peer (see kf_exec_reference_order) */
    st->cw = st->cols / st->nchunks;
    for (int j = 0; j <= st->nchunks; ++j) st->coff[j] = j * st->cw;
    /* a narrower last chunk: C of the last chunk is the one pass nothing overlaps (every B of every rank precedes it) */
    if (st->nchunks >= 2 && st->tail16 > 0 && st->tail16 < 16) {
        int tail = (int)((long long)st->cols * st->tail16 / 16) / 16 * 16;
        if (tail < 16) tail = 16;
        const int head = st->cols - tail, nh = st->nchunks - 1;
        if (head > 0 && head % (16 * nh) == 0 && tail < head / nh) {
            for (int j = 0; j < nh; ++j) st->coff[j] = j * (head / nh);
            st->coff[nh] = head;
            st->coff[st->nchunks] = st->cols;
        }
    }
}

/* testing aid (host logic only, no CUDA call): the column chunks a cfg with these parameters would use; returns their number
 * and writes the nchunks + 1 boundaries */
int kiss_fftnd_mgpu_debug_chunks(int cols, int nranks, int want, int tail16, int *coff, int cap)
{
    struct kiss_fftnd_mgpu_state tmp;
    memset(&tmp, 0, sizeof(tmp));
    tmp.cols = cols; tmp.nranks = nranks; tmp.tail16 = tail16;
    kf_set_chunks(&tmp, want);
    for (int j = 0; j <= tmp.nchunks && j < cap; ++j) coff[j] = tmp.coff[j];
    return tmp.nchunks;
}

/* Tuning aid (tests/cpp/test_mgpu.c sweeps with it): change the pipeline shape of an existing cfg.  COLLECTIVE -- every
 * rank passes the same values, with no exec in flight.  knobs[0] chunks of the k2 columns, [1] plane groups, [2] CTA cap
 * of the link-bound launches (0 = none), [3] priority of their stream (0/1); a negative entry keeps the current value. */
int kiss_fftnd_mgpu_tune(kiss_fftnd_mgpu_cfg st, const int *knobs, int nknobs)
{
    if (!st || st->magic != KF_MAGIC_MGPU || !knobs) return KISS_FFT_CUDA_EINVAL;
    CU(cudaDeviceSynchronize());
    if (nknobs > 0 && knobs[0] > 0) kf_set_chunks(st, knobs[0]);
    if (nknobs > 1 && knobs[1] > 0 && knobs[1] <= KF_MGPU_MAXCHUNKS && st->planes % knobs[1] == 0 &&
        !(st->flags & KISS_FFT_MGPU_REFERENCE_ORDER)) st->pchunks = knobs[1];
    if (nknobs > 2 && knobs[2] >= 0) st->b_ctas = knobs[2];
    if (nknobs > 3 && knobs[3] >= 0 && (knobs[3] != 0) != (st->b_prio != 0)) CU(kf_make_b_stream(st, knobs[3]));
    if (nknobs > 4 && knobs[4] >= 0) st->ac_reserve = knobs[4];
    if (nknobs > 5 && knobs[5] >= 0) st->trace = knobs[5];
    if (nknobs > 8 && knobs[8] >= 0) { st->tail16 = knobs[8]; kf_set_chunks(st, st->nchunks); }
    if (nknobs > 6 && knobs[6] >= 0 && st->p2p && knobs[6] != (st->gc_on ? st->link_sms : 0)) kf_make_partition(st, knobs[6]);
    return 0;
}
int kiss_fftnd_mgpu_knob(kiss_fftnd_mgpu_cfg st, int which)
{
    if (!st) return -1;
    return which == 0 ? st->nchunks : which == 1 ? st->pchunks : which == 2 ? st->b_ctas : which == 3 ? st->b_prio : which == 4 ? st->ac_reserve :
           which == 6 ? (st->gc_on ? st->link_sms : 0) : which == 7 ? (st->gc_on ? st->rest_sms : 0) : which == 8 ? st->coff[st->nchunks] - st->coff[st->nchunks - 1] : -1;
}

kiss_fftnd_mgpu_cfg kiss_fftnd_mgpu_alloc(const int *dims, int ndims, int inverse_fft, int rank, int nranks, const void *id,
                                          unsigned flags)
{
    kiss_fftnd_mgpu_cfg st = NULL;
    if (!dims || (ndims != 3 && ndims != 2) || nranks < 1 || nranks > KF_MGPU_MAXRANKS || rank < 0 || rank >= nranks || (nranks > 1 && !id)) {
        kf_fail(NULL, "kiss_fftnd_mgpu_alloc: bad argument (2-D or 3-D arrays, 1..16 ranks, id required beyond one rank)", KISS_FFT_CUDA_EINVAL, 0);
        return NULL;
    }
    /* 2-D n0 x n1 = the 3-D geometry n0 x 1 x n1 (slabs of n0/G rows; output X[k0][k1] stored [k1 - r*n1/G][k0]); it has
     * its own, simpler exec (kf_exec_2d): rows, transposing exchange, rows */
    int dims3[3] = {dims[0], ndims == 2 ? 1 : dims[1], ndims == 2 ? dims[1] : dims[2]};
    if (ndims == 2 && (flags & KISS_FFT_MGPU_REFERENCE_ORDER)) {
        kf_fail(NULL, "kiss_fftnd_mgpu_alloc: the reference-order mode is 3-D only", KISS_FFT_CUDA_EINVAL, 0);
        return NULL;
    }
    dims = dims3;
    if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || dims[0] % nranks || dims[2] % nranks) {
        kf_fail(NULL, "kiss_fftnd_mgpu_alloc: dims[0] and dims[2] must be divisible by the number of ranks", KISS_FFT_CUDA_EINVAL, 0);
        return NULL;
    }
    st = (kiss_fftnd_mgpu_cfg)calloc(1, sizeof(*st));
    if (!st) return NULL;
    st->magic = KF_MAGIC_MGPU;
    st->d0 = dims[0]; st->d1 = dims[1]; st->d2 = dims[2];
    st->inverse = inverse_fft ? 1 : 0;
    st->rank = rank; st->nranks = nranks; st->flags = flags;
    st->ndims = ndims;
    st->planes = st->d0 / nranks;
    st->cols = st->d2 / nranks;
    /* chunks of the k2 columns: as many as requested / up to 4, keeping whole 16-column tiles per chunk where possible */
    int want = 2;                      /* measured at G = 2, 4, 8 (profiles/r02/mgpu_*sweep*): more chunks only add launches and gaps */
    const char *env = getenv("KISSFFT_MGPU_CHUNKS");
    if (env && atoi(env) > 0) want = atoi(env);
    st->tail16 = 4;                    /* last chunk = a quarter of the columns (the others share the rest) */
    env = getenv("KISSFFT_MGPU_TAIL16");
    if (env && atoi(env) >= 0) st->tail16 = atoi(env);
    if (ndims == 2) want = 1;
    kf_set_chunks(st, want);
    st->pchunks = 1;
    if (nranks > 1 && !(flags & KISS_FFT_MGPU_REFERENCE_ORDER))
        for (int c = 2; c >= 1; --c)
            if (st->planes % c == 0) { st->pchunks = c; break; }
    env = getenv("KISSFFT_MGPU_PCHUNKS");
    if (env && atoi(env) > 0 && atoi(env) <= KF_MGPU_MAXCHUNKS && st->planes % atoi(env) == 0) st->pchunks = atoi(env);
    /* link-bound B launches need sms * (B at full rate / its NVLink time) CTAs: about 0.32 * G/(G-1) of the device
     * (2 x 8 GiB/G at 5.7 TB/s against 8 GiB (G-1)/G^2 at 0.77 TB/s), measured best at G = 2: 96 of 148 */
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        st->dev_sms = sms;
        st->b_ctas = nranks > 1 ? (int)(0.325 * sms * nranks / (nranks - 1)) : 0;
    }
    env = getenv("KISSFFT_MGPU_B_CTAS");
    if (env && atoi(env) >= 0) st->b_ctas = atoi(env);
    /* A persistent kernel that fills the device keeps every later kernel off the SMs until it ends, and one that arrives
     * while another holds part of the device runs its statically striped tiles in two waves -- so overlap between streams
     * needs the SMs partitioned by hand: the HBM-bound launches that run beside link-bound ones leave those SMs alone. */
    st->ac_reserve = (flags & KISS_FFT_MGPU_P2P) ? st->b_ctas : 0;
    env = getenv("KISSFFT_MGPU_AC_RESERVE");
    if (env && atoi(env) >= 0) st->ac_reserve = atoi(env);
    int ok = cudaGetDevice(&st->device) == cudaSuccess;
    st->cfg0 = kiss_fft_alloc(st->d0, st->inverse, NULL, NULL);
    st->cfg1 = ndims == 2 ? NULL : kiss_fft_alloc(st->d1, st->inverse, NULL, NULL);
    st->cfg2 = kiss_fft_alloc(st->d2, st->inverse, NULL, NULL);
    ok = ok && st->cfg0 && (st->cfg1 || ndims == 2) && st->cfg2;
    st->recv_bytes = sizeof(kiss_fft_cpx) * (size_t)st->d0 * st->cols * st->d1;
    st->recv_bytes = (st->recv_bytes + 255u) & ~(size_t)255u;
    ok = ok && cudaMalloc((void **)&st->recv_base, st->recv_bytes + KF_FLAG_BYTES) == cudaSuccess &&
         cudaMemset(st->recv_base + st->recv_bytes, 0, KF_FLAG_BYTES) == cudaSuccess;
    if ((flags & KISS_FFT_MGPU_REFERENCE_ORDER) || (ndims == 2 && nranks > 1 && !(flags & KISS_FFT_MGPU_P2P)))
        ok = ok && cudaMalloc((void **)&st->work, st->recv_bytes) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&st->s_comm, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&st->s_c, cudaStreamNonBlocking) == cudaSuccess &&
         kf_make_b_stream(st, getenv("KISSFFT_MGPU_PRIO") ? atoi(getenv("KISSFFT_MGPU_PRIO")) : 1) == 0 &&
         cudaEventCreateWithFlags(&st->ev_start, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&st->ev_bdone, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&st->ev_done, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&st->ev_done2, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&st->ev_adone, cudaEventDisableTiming) == cudaSuccess;
    for (int j = 0; ok && j < KF_MGPU_MAXCHUNKS; ++j)
        ok = cudaEventCreateWithFlags(&st->ev_b[j], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st->ev_a[j], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&st->ev_x[j], cudaEventDisableTiming) == cudaSuccess;
    if (ok && nranks > 1) {
        ok = kf_nccl_load() == 0;
        if (!ok) kf_fail(st, "libnccl.so.2 could not be loaded", KISS_FFT_CUDA_EINVAL, 0);
        if (ok) {
            kf_nccl_id nid;
            memcpy(&nid, id, sizeof(nid));
            const int rc = g_nccl.CommInitRank(&st->comm, nranks, nid, rank);
            if (rc != KF_NCCL_SUCCESS) { kf_fail(st, "ncclCommInitRank", rc, 1); ok = 0; }
        }
        if (ok && (flags & KISS_FFT_MGPU_P2P)) st->p2p = kf_mgpu_map_peers(st) == 0;   /* falls back to NCCL when IPC is unavailable */
        if (ok && !st->p2p) ok = cudaMalloc((void **)&st->send, st->recv_bytes) == cudaSuccess;
        if (ok && !st->p2p && ndims == 2 && !st->work) ok = cudaMalloc((void **)&st->work, st->recv_bytes) == cudaSuccess;   /* IPC fell back to NCCL */
        if (ok && st->p2p && ndims == 3 && !(flags & KISS_FFT_MGPU_REFERENCE_ORDER)) {
            /* SMs of the link partition.  B must keep NVLink busy (0.62-0.65 TB/s measured for SM stores at every G): at
             * G = 2 it moves as many HBM bytes as link bytes and needs half the device (72 = one die), at G >= 4 it is
             * link-bound.  Measured on 1024^3 (profiles/r02/mgpu_*_green_context_sweep*): G = 2: 72, G = 4: 48, G = 8: 64. */
            int want_link = nranks <= 2 ? 72 : nranks <= 4 ? 48 : 64;
            if (want_link + 16 > st->dev_sms) want_link = st->dev_sms / 2 / 8 * 8;
            env = getenv("KISSFFT_MGPU_LINK_SMS");
            if (env && atoi(env) >= 0) want_link = atoi(env);
            kf_make_partition(st, want_link);
        }
    }
    if (ok) ok = cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {
        if (!st->err[0]) kf_fail(st, "kiss_fftnd_mgpu_alloc: resource allocation failed", (int)cudaGetLastError(), 0);
        kiss_fftnd_mgpu_free(st);
        return NULL;
    }
    st->peer_base[rank] = st->recv_base;
    return st;
}

void kiss_fftnd_mgpu_free(kiss_fftnd_mgpu_cfg st)
{
    if (!st || st->magic != KF_MAGIC_MGPU) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < st->nranks; ++r)
        if (r != st->rank && st->peer_base[r]) cudaIpcCloseMemHandle(st->peer_base[r]);
    if (st->d_peer_flags) cudaFree(st->d_peer_flags);
    if (st->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(st->comm);
    if (st->recv_base) cudaFree(st->recv_base);
    if (st->send) cudaFree(st->send);
    if (st->work) cudaFree(st->work);
    if (st->gc_on || st->gs_b) kf_drop_partition(st);
    if (st->ev_done2) cudaEventDestroy(st->ev_done2);
    if (st->ev_adone) cudaEventDestroy(st->ev_adone);
    if (st->s_comm) cudaStreamDestroy(st->s_comm);
    if (st->s_c) cudaStreamDestroy(st->s_c);
    if (st->s_b) cudaStreamDestroy(st->s_b);
    if (st->ev_bdone) cudaEventDestroy(st->ev_bdone);
    if (st->ev_start) cudaEventDestroy(st->ev_start);
    if (st->ev_done) cudaEventDestroy(st->ev_done);
    for (int j = 0; j < KF_MGPU_MAXCHUNKS; ++j) {
        if (st->ev_b[j]) cudaEventDestroy(st->ev_b[j]);
        if (st->ev_a[j]) cudaEventDestroy(st->ev_a[j]);
        if (st->ev_x[j]) cudaEventDestroy(st->ev_x[j]);
    }
    for (int n = 0; n < KF_MGPU_MAXTRACE; ++n) {
        if (st->tr_ev[n][0]) cudaEventDestroy(st->tr_ev[n][0]);
        if (st->tr_ev[n][1]) cudaEventDestroy(st->tr_ev[n][1]);
    }
    if (st->tr_t0) cudaEventDestroy(st->tr_t0);
    kiss_fft_free(st->cfg0);
    if (st->cfg1) kiss_fft_free(st->cfg1);
    kiss_fft_free(st->cfg2);
    st->magic = 0;
    free(st);
}

/* both layouts hold d0*d1*d2/G elements per rank: [P][d1][d2] / [C][d1][d0] (fast order), [d0][d1][C] / [P][d1][d2] (reference order) */
size_t kiss_fftnd_mgpu_local_in_elems(kiss_fftnd_mgpu_cfg st) { return st ? (size_t)st->planes * st->d1 * st->d2 : 0; }
size_t kiss_fftnd_mgpu_local_out_elems(kiss_fftnd_mgpu_cfg st) { return st ? (size_t)st->cols * st->d1 * st->d0 : 0; }
int kiss_fftnd_mgpu_uses_p2p(kiss_fftnd_mgpu_cfg st) { return st ? st->p2p : 0; }
int kiss_fftnd_mgpu_chunks(kiss_fftnd_mgpu_cfg st) { return st ? st->nchunks : 0; }
size_t kiss_fftnd_mgpu_a2a_bytes(kiss_fftnd_mgpu_cfg st)
{
    return st ? sizeof(kiss_fft_cpx) * (size_t)(st->nranks - 1) * st->planes * st->cols * st->d1 : 0;
}

/* The reference's axis order 0, 1, 2 across GPUs (KISS_FFT_MGPU_REFERENCE_ORDER; SURVEY.md 8e "fixed-point N-D").
 * kiss_fftnd.c:156-188 sweeps the axes in the order 0, 1, ..., and the Q15 / Q31 roundings depend on that order, so the
 * fast pipeline above (axes 2, 1, 0) is not bit-identical to it.  Starting from slabs along the LAST axis -- rank r holds
 * x[i0][i1][r*C + c] as [d0][d1][C] -- axes 0 and 1 are local, the same single all-to-all completes axis 2, and the result
 * comes out in natural order as axis-0 slabs: rank r holds X[r*P + p][k1][k2] as [P][d1][d2], the very rows kiss_fftnd
 * would have written.  Every pass is the transposing axis pass of kiss_fftnd.c:172-178 on the reference's operands:
 *   1. axis 0:  [d0][d1*C]  ->  work [d1][C][d0]                                     kiss_fft_axis_pass_dev
 *   2. axis 1:  work viewed as C planes of d0 columns (stride C*d0)  ->  [C][d0][d1], column block s = planes s*P.. of
 *      the result goes to rank s, which collects [d2 = r*C + c][P][d1]              kiss_fft_planes_pass_peers2_dev
 *   3. axis 2:  [d2][P*d1]  ->  d_out [P][d1][d2]                                    kiss_fft_axis_pass_dev
 * Same butterflies on the same operands in the same order => bit-identical to the single-GPU kiss_fftnd in every
 * datatype.  One exchange block per peer (no chunk pipeline: this is the exact mode, not the fast one). */
static int kf_exec_reference_order(kiss_fftnd_mgpu_cfg st, kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, cudaStream_t main)
{
    const int G = st->nranks, P = st->planes, d0 = st->d0, d1 = st->d1, C = st->cols;
    kiss_fft_cpx *recv = (kiss_fft_cpx *)st->recv_base;
    const size_t blk = (size_t)C * P * d1;                 /* what one rank sends to one rank */
    const unsigned epoch = ++st->epoch;
    CU(kiss_fft_axis_pass_dev(st->cfg0, d_in, st->work, (size_t)d1 * C, (size_t)d1 * C, main));
    kiss_fft_cpx *dst[KF_MGPU_MAXRANKS];
    for (int s = 0; s < G; ++s)
        dst[s] = (G > 1 && !st->p2p) ? st->send + (size_t)s * blk : (kiss_fft_cpx *)st->peer_base[s] + (size_t)st->rank * blk;
    if (G > 1 && st->p2p) {
        /* every peer has finished reading its receive buffer (step 3 of the previous call) before anyone stores into it */
        CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, 1, epoch, main));
        CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, 1, epoch, main));
    }
    CU(kiss_fft_planes_pass_peers2_dev(st->cfg1, st->work, (kiss_fft_cpx *const *)dst, G, (size_t)C, (size_t)P, (size_t)P,
                                       (size_t)C * d0, (size_t)d0, (size_t)P * d1, 0, 0, main));
    if (G > 1 && st->p2p) {
        CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, 0, epoch, main));
        CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, 0, epoch, main));
    } else if (G > 1) {
        NC(g_nccl.GroupStart());
        for (int s = 0; s < G; ++s) {
            NC(g_nccl.Send(st->send + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, main));
            NC(g_nccl.Recv(recv + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, main));
        }
        NC(g_nccl.GroupEnd());
    }
    CU(kiss_fft_axis_pass_dev(st->cfg2, recv, d_out, (size_t)P * d1, (size_t)P * d1, main));
    return 0;
}

/* 2-D arrays: rows (axis 1) in place -> tiled transposing exchange (every destination row piece is P contiguous elements,
 * stored straight into the owner's receive buffer [C][d0], or through a send buffer and NCCL) -> rows (axis 0) of what
 * arrived.  kiss_fftnd.c:156-188 on two axes with the transposing write of its first sweep done by the exchange. */
static int kf_exec_2d(kiss_fftnd_mgpu_cfg st, kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, cudaStream_t main)
{
    const int G = st->nranks, P = st->planes, d0 = st->d0, d2 = st->d2, C = st->cols;
    kiss_fft_cpx *recv = (kiss_fft_cpx *)st->recv_base;
    void *dst[KF_MGPU_MAXRANKS];
    CU(kiss_fft_batch_dev(st->cfg2, d_in, d_in, (size_t)P, (size_t)d2, (size_t)d2, 1, main));
    if (G == 1) {
        dst[0] = recv;
        CU(kfcu_transpose_peers(d_in, d2, dst, 1, P, C, d0, 0, main));
    } else if (st->p2p) {
        const unsigned epoch = ++st->epoch;
        /* every peer has finished reading its receive buffer (the rows pass of the previous call) before anyone stores into it */
        CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, 1, epoch, main));
        CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, 1, epoch, main));
        for (int s = 0; s < G; ++s) dst[s] = st->peer_base[s];
        CU(kfcu_transpose_peers(d_in, d2, dst, G, P, C, d0, (long long)st->rank * P, main));
        CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, 0, epoch, main));
        CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, 0, epoch, main));
    } else {
        const size_t blk = (size_t)C * P;            /* send [s][c][p]; received as work [r][c][p] */
        for (int s = 0; s < G; ++s) dst[s] = st->send + (size_t)s * blk;
        CU(kfcu_transpose_peers(d_in, d2, dst, G, P, C, P, 0, main));
        NC(g_nccl.GroupStart());
        for (int s = 0; s < G; ++s) {
            NC(g_nccl.Send(st->send + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, main));
            NC(g_nccl.Recv(st->work + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, main));
        }
        NC(g_nccl.GroupEnd());
        for (int r = 0; r < G; ++r)                  /* [r][c][p] -> [c][r*P + p] */
            CU(cudaMemcpy2DAsync(recv + (size_t)r * P, sizeof(kiss_fft_cpx) * (size_t)d0, st->work + (size_t)r * blk, sizeof(kiss_fft_cpx) * (size_t)P,
                                 sizeof(kiss_fft_cpx) * (size_t)P, (size_t)C, cudaMemcpyDeviceToDevice, main));
    }
    CU(kiss_fft_batch_dev(st->cfg0, recv, d_out, (size_t)C, (size_t)d0, (size_t)d0, 1, main));
    return 0;
}

/* ---- tuning aid: a timeline of the launches of one exec ------------------------------------------------------- */
static void tr_begin(kiss_fftnd_mgpu_cfg st, const char *what, int i, int j, cudaStream_t s)
{
    if (!st->trace || st->ntr >= KF_MGPU_MAXTRACE) return;
    const int n = st->ntr;
    if (!st->tr_ev[n][0]) { cudaEventCreate(&st->tr_ev[n][0]); cudaEventCreate(&st->tr_ev[n][1]); }
    snprintf(st->tr_name[n], sizeof(st->tr_name[n]), "%s%d.%d", what, i, j);
    cudaEventRecord(st->tr_ev[n][0], s);
}
static void tr_end(kiss_fftnd_mgpu_cfg st, cudaStream_t s)
{
    if (!st->trace || st->ntr >= KF_MGPU_MAXTRACE) return;
    cudaEventRecord(st->tr_ev[st->ntr++][1], s);
}
/* after the exec has completed: writes "name start_ms end_ms" lines (relative to the start of the exec) into buf */
int kiss_fftnd_mgpu_trace(kiss_fftnd_mgpu_cfg st, char *buf, size_t len)
{
    if (!st || st->magic != KF_MAGIC_MGPU || !buf || !len) return KISS_FFT_CUDA_EINVAL;
    size_t o = 0;
    buf[0] = 0;
    CU(cudaDeviceSynchronize());
    for (int n = 0; n < st->ntr && o + 48 < len; ++n) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, st->tr_t0, st->tr_ev[n][0]);
        cudaEventElapsedTime(&b, st->tr_t0, st->tr_ev[n][1]);
        o += (size_t)snprintf(buf + o, len - o, "%s %.3f %.3f\n", st->tr_name[n], a, b);
    }
    return 0;
}

int kiss_fftnd_mgpu_exec(kiss_fftnd_mgpu_cfg st, kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, void *stream)
{
    if (!st || st->magic != KF_MAGIC_MGPU || !d_in || !d_out) return kf_fail(st, "kiss_fftnd_mgpu_exec: bad argument", KISS_FFT_CUDA_EINVAL, 0);
    if (st->flags & KISS_FFT_MGPU_REFERENCE_ORDER) return kf_exec_reference_order(st, d_in, d_out, (cudaStream_t)stream);
    if (st->ndims == 2) return kf_exec_2d(st, d_in, d_out, (cudaStream_t)stream);
    const int G = st->nranks, P = st->planes, d0 = st->d0, d1 = st->d1, d2 = st->d2, C = st->cols;
    cudaStream_t main = (cudaStream_t)stream;
    kiss_fft_cpx *recv = (kiss_fft_cpx *)st->recv_base;
    if (G == 1) {
        /* one rank: rows, then B writes the whole planes transposed into the "receive" buffer, C follows -- no exchange */
        CU(kiss_fft_batch_dev(st->cfg2, d_in, d_in, (size_t)P * d1, (size_t)d2, (size_t)d2, 1, main));
        CU(kiss_fft_planes_pass_dev(st->cfg1, d_in, recv, (size_t)P, (size_t)d2, (size_t)d2, (size_t)d1 * d2, (size_t)d2 * d1, main));
        CU(kiss_fft_axis_pass_dev(st->cfg0, recv, d_out, (size_t)C * d1, (size_t)C * d1, main));
        return 0;
    }
    const unsigned epoch = ++st->epoch;
    const int NP = st->pchunks, Pc = P / NP;
    if (st->trace) {
        if (!st->tr_t0) CU(cudaEventCreate(&st->tr_t0));
        st->ntr = 0;
        CU(cudaEventRecord(st->tr_t0, main));
    }
    /* streams: with an SM partition B runs in the link partition, A(i > 0) and C(j < last) in the other one */
    const int gc = st->gc_on && st->p2p;
    cudaStream_t sB = gc ? st->gs_b : st->s_b, sA = gc ? st->gs_a : main, sC = gc ? st->gs_c : st->s_c;
    const int rsv_rest = gc ? st->dev_sms - st->rest_sms : st->ac_reserve;      /* SMs the HBM-bound launches leave alone */
    CU(cudaEventRecord(st->ev_start, main));
    CU(cudaStreamWaitEvent(st->s_c, st->ev_start, 0));
    CU(cudaStreamWaitEvent(sB, st->ev_start, 0));
    if (gc) CU(cudaStreamWaitEvent(sC, st->ev_start, 0));
    if (st->p2p) {
        /* every peer has finished reading its receive buffer (step C of the previous call) before anyone stores into it */
        CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, st->nchunks, epoch, sB));
        CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, st->nchunks, epoch, sB));
    } else {
        CU(cudaStreamWaitEvent(st->s_comm, st->ev_start, 0));
    }
    /* A(i): rows of plane group i, in place; A(0) on the caller's stream and the whole device, the later groups beside B */
    for (int i = 0; i < NP; ++i) {
        kiss_fft_cpx *rows = d_in + (size_t)i * Pc * d1 * d2;
        cudaStream_t sa = i > 0 ? sA : main;
        if (i == 1 && sa != main) CU(cudaStreamWaitEvent(sa, st->ev_a[0], 0));
        kfcu_set_sm_reserve(i > 0 ? rsv_rest : 0);
        tr_begin(st, "A", i, 0, sa);
        const int rc_a = kiss_fft_batch_dev(st->cfg2, rows, rows, (size_t)Pc * d1, (size_t)d2, (size_t)d2, 1, sa);
        tr_end(st, sa);
        kfcu_set_sm_reserve(0);
        CU(rc_a);
        CU(cudaEventRecord(st->ev_a[i], sa));
    }
    for (int j = 0; j < st->nchunks; ++j) {
        const int c0 = st->coff[j], cwj = st->coff[j + 1] - c0;              /* columns [c0, c0 + cwj) of every rank's range */
        const size_t cbase = (size_t)c0 * d0 * d1, blk = (size_t)P * cwj * d1; /* chunk j in the receive / send buffer; one (source, chunk) block */
        for (int i = 0; i < NP; ++i) {
            /* B(i, j): k2 columns s*C + c0 + [0, cwj) of the planes of group i go to rank s, transposed */
            kiss_fft_cpx *dst[KF_MGPU_MAXRANKS];
            for (int s = 0; s < G; ++s) {
                if (st->p2p)    /* the receiver's layout [chunk][c][i0 = me*P + plane][d1] */
                    dst[s] = (kiss_fft_cpx *)st->peer_base[s] + cbase + ((size_t)st->rank * P + (size_t)i * Pc) * d1;
                else            /* the local send buffer [chunk][dest s][P][cw][d1] */
                    dst[s] = st->send + cbase + (size_t)s * blk + (size_t)i * Pc * cwj * d1;
            }
            if (j == 0) CU(cudaStreamWaitEvent(sB, st->ev_a[i], 0));
            /* partition: the grid is sized to the link partition.  NCCL path: B is HBM-bound itself; beside a running
             * exchange (j > 0) it leaves the SMs NCCL's kernel needs */
            kfcu_set_sm_reserve(gc ? st->dev_sms - st->link_sms : (!st->p2p && j > 0 ? st->ac_reserve : 0));
            tr_begin(st, "B", i, j, sB);
            const int rc_b = kiss_fft_planes_pass_peers2_dev(st->cfg1, d_in + (size_t)i * Pc * d1 * d2 + (size_t)c0, (kiss_fft_cpx *const *)dst, G,
                                                             (size_t)Pc, (size_t)cwj, (size_t)C, (size_t)d2, (size_t)d1 * d2,
                                                             st->p2p ? (size_t)d1 : (size_t)cwj * d1, st->p2p ? (size_t)d0 * d1 : 0,
                                                             st->p2p && !gc ? st->b_ctas : 0, sB);
            tr_end(st, sB);
            kfcu_set_sm_reserve(0);
            CU(rc_b);
        }
        const int last = j + 1 == st->nchunks;
        cudaStream_t sc = last ? st->s_c : sC;      /* the last C runs after every B of every rank: whole device */
        if (st->p2p) {
            CU(kfcu_peer_signal((void *const *)st->d_peer_flags, G, st->rank, j, epoch, sB));
            /* the spinning wait is submitted only once the local stores of this chunk are done (the peers finish theirs at
             * about the same time): a kernel that spins from the start of the exec at the head of a hardware queue holds
             * up whatever other stream shares that queue (CUDA_DEVICE_MAX_CONNECTIONS) */
            CU(cudaEventRecord(st->ev_b[j], sB));
            CU(cudaStreamWaitEvent(sc, st->ev_b[j], 0));
            CU(kfcu_peer_wait(flag_ptr(st, st->rank), G, j, epoch, sc));
        } else {
            CU(cudaEventRecord(st->ev_b[j], sB));
            CU(cudaStreamWaitEvent(st->s_comm, st->ev_b[j], 0));
            tr_begin(st, "X", 0, j, st->s_comm);
            NC(g_nccl.GroupStart());
            for (int s = 0; s < G; ++s) {
                NC(g_nccl.Send(st->send + cbase + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, st->s_comm));
                NC(g_nccl.Recv(recv + cbase + (size_t)s * blk, blk * sizeof(kiss_fft_cpx), KF_NCCL_UINT8, s, st->comm, st->s_comm));
            }
            NC(g_nccl.GroupEnd());
            tr_end(st, st->s_comm);
            CU(cudaEventRecord(st->ev_x[j], st->s_comm));
            CU(cudaStreamWaitEvent(sc, st->ev_x[j], 0));
        }
        /* C(j): axis 0 of what has arrived from every rank */
        int rc_c;
        tr_begin(st, "C", 0, j, sc);
        kfcu_set_sm_reserve(last ? 0 : rsv_rest);
        if (st->p2p)
            rc_c = kiss_fft_planes_pass_dev(st->cfg0, recv + cbase, d_out + (size_t)c0 * d1 * d0, (size_t)cwj, (size_t)d1, (size_t)d1,
                                            (size_t)d0 * d1, (size_t)d1 * d0, sc);
        else
            rc_c = kiss_fft_axis_pass_dev(st->cfg0, recv + cbase, d_out + (size_t)c0 * d1 * d0, (size_t)cwj * d1, (size_t)cwj * d1, sc);
        tr_end(st, sc);
        kfcu_set_sm_reserve(0);
        CU(rc_c);
    }
    CU(cudaEventRecord(st->ev_bdone, sB));
    CU(cudaEventRecord(st->ev_done, st->s_c));
    CU(cudaStreamWaitEvent(main, st->ev_bdone, 0));
    CU(cudaStreamWaitEvent(main, st->ev_done, 0));
    if (gc) {
        CU(cudaEventRecord(st->ev_done2, sC));
        CU(cudaStreamWaitEvent(main, st->ev_done2, 0));
        if (NP > 1) {       /* implied by the B launches that waited for every A, stated for the reader */
            CU(cudaEventRecord(st->ev_adone, sA));
            CU(cudaStreamWaitEvent(main, st->ev_adone, 0));
        }
    }
    return 0;
}
