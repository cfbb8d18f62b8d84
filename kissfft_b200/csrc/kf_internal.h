/* kf_internal.h -- private C interface between the C host layer (kf_api.c) and the CUDA launchers
 * (kf_launch.cu).  One library is built per datatype, exactly like the reference builds one
 * libkissfft-<type>.so per datatype (reference Makefile:89-137); the datatype is fixed by the same macros the
 * reference uses (FIXED_POINT=16|32, kiss_fft_scalar=float|double). */
#ifndef KF_INTERNAL_H
#define KF_INTERNAL_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KFCU_MAXSTAGES 32 /* MAXFACTORS, _kiss_fft_guts.h:21 */

enum { KFCU_C2C = 0, KFCU_C2C_COL = 1, KFCU_R2C = 2, KFCU_C2R = 3 };

/* device-side view of one 1-D plan (one per (device, nfft, inverse)) */
typedef struct kfcu_plan {
    int nfft; /* length of the complex transform (ncfft for the real modes) */
    int inverse;
    int nstages;
    int p[KFCU_MAXSTAGES]; /* kf_factor order: p[0] outermost */
    int m[KFCU_MAXSTAGES];
    const void *d_tw;  /* nfft complex twiddles in device memory */
    const void *d_stw; /* nfft/2 split twiddles (real modes) or NULL */
    const void *h_tw;  /* the same twiddles on the host (for the butterfly constants) */
    void *d_gtw[5];    /* per-group stage-twiddle tables of the fused plan serving each mode (KFCU_C2C..KFCU_C2R; [4]: the
                          fast-convolution plan);
                          created lazily by kf_launch.cu, freed by kiss_fft_cleanup together with d_tw */
} kfcu_plan;

/* Runs `howmany` transforms. Distances are in complex elements of the respective side (for KFCU_R2C the input
 * side and for KFCU_C2R the output side are real rows viewed as packed complex, i.e. scalars/2).
 * Returns 0 or a cudaError_t value; KFCU_E* (negative) for argument errors. */
int kfcu_exec(int mode, kfcu_plan *plan, const void *d_in, void *d_out, long long howmany, long long in_dist,
              long long out_dist, long long in_stride, void *stream);

/* KFCU_C2C_COL over `nplanes` independent planes (slab decomposition): plane p, column c reads
 * in[p*in_pdist + c + j*col_stride], j < nfft, and writes out[p*out_pdist + c*nfft + k] */
int kfcu_exec_planes(kfcu_plan *plan, const void *d_in, void *d_out, long long nplanes, long long ncols, long long col_stride,
                     long long in_pdist, long long out_pdist, void *stream);

/* Four-step transform of nrows rows of length N = nfft*ncols (float / double), one pass per call:
 *   step 0: plan of length N1 = nfft, ncols = N2: column n2 of a row viewed as [N1][N2] -> out[row*N + n2*N1 + k1] * W_N^(n2*k1)
 *           (d_twbig: the N twiddles of the long transform);
 *   step 1: plan of length N2 = nfft, ncols = N1: column k1 of the intermediate array viewed as [N2][N1] ->
 *           out[row*N + k2*N1 + k1], natural order.
 * kfcu_has_fourstep(nfft): both passes are available for this length. */
int kfcu_exec_fourstep(kfcu_plan *plan, int step, const void *d_in, void *d_out, long long nrows, long long ncols,
                       const void *d_twbig, void *stream);
int kfcu_has_fourstep(int nfft);
int kfcu_has_colcol(int nfft);   /* step 1 alone (every datatype): an axis pass that keeps the array layout */
int kfcu_has_colring(int nfft);  /* the transposing column pass of this length has a tensor-map input-ring variant */

/* kfcu_exec_planes with the ncols = npeers*cols_per_peer columns of every plane scattered to npeers destination
 * buffers: column block s is written through peers[s] (+ p*out_pdist + c_local*nfft) */
int kfcu_exec_planes_peers(kfcu_plan *plan, const void *d_in, void *const *peers, int npeers, long long nplanes,
                           long long cols_per_peer, long long peer_col_dist, long long col_stride, long long in_pdist,
                           long long out_pdist, long long out_col_dist, int max_ctas, void *stream);
/* flags between the GPUs of the slab transform: d_flag_ptrs = device array of every rank's flag words (mapped here); a
 * signal sets word (slot*16 + rank) of every rank to `epoch`, a wait spins until all nranks words of `slot` reached it */
int kfcu_peer_signal(void *const *d_flag_ptrs, int nranks, int rank, int slot, unsigned epoch, void *stream);
int kfcu_peer_wait(const void *d_my_flags, int nranks, int slot, unsigned epoch, void *stream);

/* Multi-pass path: stage s of the plan as one launch over global memory (levels in autosort layout, dense rows of nfft
 * in the work buffers).  first: read the caller's rows (in_dist / in_stride); last: write the caller's rows (out_dist). */
int kfcu_stage(const kfcu_plan *plan, int s, const void *d_in, void *d_out, long long batch, long long in_dist,
               long long out_dist, long long in_stride, int first, int last, void *stream);
/* stand-alone split pass of the real transforms: post != 0: T[nc] -> F[nc+1] (kiss_fftr.c:88-116), else F -> T */
int kfcu_realpass(const kfcu_plan *plan, int post, const void *d_in, void *d_out, long long batch, long long in_dist,
                  long long out_dist, void *stream);

/* fused overlap-scrap fast convolution (float / double builds): block b reads nfft samples at d_in + b*ngood, writes
 * ngood samples at d_out + b*ngood; d_h = nfft-point frequency response already scaled by 1/nfft.  KFCU_ETOOBIG when
 * no fused plan exists for the length (kf_api.c then composes it from the pieces below: kf_fastconv_unfused). */
int kfcu_has_fastconv(int nfft);
int kfcu_fastconv(kfcu_plan *fwd, kfcu_plan *inv, const void *d_in, void *d_out, long long nblocks, long long ngood,
                  const void *d_h, void *stream);

/* unfused fast convolution: dense copies of the overlapping blocks (out[b][i] = in[b*advance + i], elements = scalars when
 * is_real else complex) and the pointwise product with the frequency response */
int kfcu_gather_blocks(const void *d_in, void *d_out, long long nblocks, int len, long long advance, int is_real, void *stream);
int kfcu_cmul_rows(void *d_x, const void *d_h, long long rows, int n, void *stream);

/* out[c][r] = in[r][c] for a rows x cols array of complex elements (kiss_fftndr's bin-major <-> row-major
 * scatter loops, kiss_fftndr.c:101-102, 107-108) */
int kfcu_transpose(const void *d_in, void *d_out, long long rows, long long cols, void *stream);
/* columns split into npeers blocks; block s lands transposed in peers[s] as [cols_per_peer][out_pitch] + out_off */
int kfcu_transpose_peers(const void *d_in, long long in_pitch, void *const *peers, int npeers, long long rows, long long cols_per_peer,
                         long long out_pitch, long long out_off, void *stream);

/* 1 if a compile-time (fused, register-group) plan exists for this length/mode in this datatype build */
int kfcu_has_fused(int nfft, int mode);
/* largest nfft the run-time (generic shared-memory) kernel accepts */
int kfcu_generic_max_nfft(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long kfcu_launch_count(void);
/* force the generic kernel even when a fused plan exists (testing aid; 0 = default) */
void kfcu_force_generic(int on);
/* cap the persistent grid of the fused kernels launched after this call (0 = no cap): lets a link-bound launch
 * (peer-memory stores of the slab exchange) share the SMs with an HBM-bound one on another stream */
void kfcu_set_grid_limit(int max_ctas);    /* launches of the calling host thread only */
void kfcu_set_sm_reserve(int sms);         /* persistent grids of the calling thread leave `sms` SMs to a concurrent kernel */

#define KFCU_EINVAL (-1)
#define KFCU_ETOOBIG (-2)

#ifdef __cplusplus
}
#endif
#endif
