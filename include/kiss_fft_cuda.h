/*
 * kiss_fft_cuda.h -- batched / device-pointer extension of the kissfft API (new in kissfft-b200).
 *
 * The reference has no batched or device interface; its callers loop kiss_fft()/kiss_fftr() over rows
 * (test/benchkiss.c:83-115, tools/fftutil.c:19-126).  These entry points are that loop, executed by one kernel
 * launch: they take the same cfg objects the reference-compatible allocators return.
 *
 * Conventions
 *   - `*_dev` functions take CUDA device pointers and a cudaStream_t (passed as void*, NULL = default stream);
 *     they enqueue work and return without synchronising.
 *   - functions without the suffix take HOST pointers, stage through pinned memory with copy/compute overlap,
 *     and return when the result is in the output buffer.
 *   - distances (`*_dist`) are in elements of the respective side (kiss_fft_cpx for complex rows,
 *     kiss_fft_scalar for real rows); real-row distances and real device pointers must be even / 2*sizeof(scalar)
 *     aligned because a real row is read as packed complex (kiss_fftr.c:77).
 *   - return value: 0 on success, a positive cudaError_t value, or a negative KISS_FFT_CUDA_E* code.  The
 *     message of the last failure in this thread is available from kiss_fft_cuda_last_error().
 */
#ifndef KISS_FFT_CUDA_H
#define KISS_FFT_CUDA_H

#include "kiss_fft.h"
#include "kiss_fftnd.h"
#include "kiss_fftndr.h"
#include "kiss_fftr.h"

#ifdef __cplusplus
extern "C" {
#endif

#define KISS_FFT_CUDA_EINVAL (-1)  /* bad argument (NULL cfg, wrong direction, odd real distance, ...) */
#define KISS_FFT_CUDA_ETOOBIG (-2) /* transform length not supported by any kernel of this build */
#define KISS_FFT_CUDA_ENOMEM (-3)

/* ---- device-pointer, stream-ordered ---------------------------------------------------------------------- */

/* howmany x kiss_fft_stride(cfg, d_in + b*in_dist, d_out + b*out_dist, in_stride)   (kiss_fft.c:375-399) */
int KISS_FFT_API kiss_fft_batch_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t howmany,
                                    size_t in_dist, size_t out_dist, int in_stride, void *stream);

/* howmany x kiss_fftr(cfg, d_time + b*time_dist, d_freq + b*freq_dist)              (kiss_fftr.c:63-117) */
int KISS_FFT_API kiss_fftr_batch_dev(kiss_fftr_cfg cfg, const kiss_fft_scalar *d_time, kiss_fft_cpx *d_freq,
                                     size_t howmany, size_t time_dist, size_t freq_dist, void *stream);

/* howmany x kiss_fftri(cfg, d_freq + b*freq_dist, d_time + b*time_dist)             (kiss_fftr.c:119-155) */
int KISS_FFT_API kiss_fftri_batch_dev(kiss_fftr_cfg cfg, const kiss_fft_cpx *d_freq, kiss_fft_scalar *d_time,
                                      size_t howmany, size_t freq_dist, size_t time_dist, void *stream);

/* kiss_fftnd on device buffers (kiss_fftnd.c:156-188).  d_work: scratch of prod(dims) complex elements or NULL
 * (then an internal scratch buffer is used).  d_in == d_out is allowed; d_in is never modified otherwise. */
int KISS_FFT_API kiss_fftnd_dev(kiss_fftnd_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, kiss_fft_cpx *d_work,
                                void *stream);

/* a single axis pass of kiss_fftnd (kiss_fftnd.c:172-178): view d_in as nfft x ncols, transform every column,
 * store column i as row i of d_out (ncols x nfft).  Building block of the multi-GPU slab transform. */
int KISS_FFT_API kiss_fft_axis_pass_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t ncols,
                                        size_t col_stride, void *stream);

/* the same axis pass over `nplanes` independent planes in one launch: plane p, column c reads
 * d_in[p*in_plane_dist + c + j*col_stride] (j < nfft) and writes row d_out[p*out_plane_dist + c*nfft + k].
 * Building block of the slab-decomposed multi-GPU 3-D transform (kissfft_b200/slab.py). */
int KISS_FFT_API kiss_fft_planes_pass_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t nplanes,
                                          size_t ncols, size_t col_stride, size_t in_plane_dist, size_t out_plane_dist,
                                          void *stream);

/* the fused "column pass + all-to-all" of the slab transform: as kiss_fft_planes_pass_dev with
 * ncols = npeers*cols_per_peer, but column block s of every plane is written through d_peers[s] -- pointers into the
 * receive buffers of the other GPUs mapped into this process (CUDA peer / symmetric memory), so the transposed rows
 * travel over NVLink as the kernel produces them: row (plane p, local column c) lands at
 * d_peers[s][p*out_plane_dist + c*nfft + k].  npeers <= 16. */
int KISS_FFT_API kiss_fft_planes_pass_peers_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *const *d_peers,
                                                int npeers, size_t nplanes, size_t cols_per_peer, size_t col_stride,
                                                size_t in_plane_dist, size_t out_plane_dist, void *stream);

/* kiss_fftndr / kiss_fftndri on device buffers (kiss_fftndr.c:86-132) */
int KISS_FFT_API kiss_fftndr_dev(kiss_fftndr_cfg cfg, const kiss_fft_scalar *d_time, kiss_fft_cpx *d_freq, void *stream);
int KISS_FFT_API kiss_fftndri_dev(kiss_fftndr_cfg cfg, const kiss_fft_cpx *d_freq, kiss_fft_scalar *d_time, void *stream);

/* ---- host-pointer batched (pinned staging + H2D / kernel / D2H overlap) ----------------------------------- */
int KISS_FFT_API kiss_fft_batch(kiss_fft_cfg cfg, const kiss_fft_cpx *in, kiss_fft_cpx *out, size_t howmany);
int KISS_FFT_API kiss_fftr_batch(kiss_fftr_cfg cfg, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata, size_t howmany);
int KISS_FFT_API kiss_fftri_batch(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata, size_t howmany);

/* ---- fused fast convolution (float / double builds only, like the reference: kiss_fastfir.c:152) ------------------
 * The overlap-scrap FIR filter of the reference's tools/kiss_fastfir.c (complex-sample build): FFT -> multiply by the
 * filter's frequency response -> IFFT for every block, fused into ONE kernel (the spectrum never leaves the SM).
 *   kiss_fastconv_alloc   kiss_fastfir_alloc (kiss_fastfir.c:59-165): *pnfft == 0 picks the reference's default size
 *   kiss_fastconv_dev     kff_nocopy / fastconv1buf (kiss_fastfir.c:167-206) on device buffers: processes every complete
 *                         nfft-sample block of the n input samples, block b reading d_in + b*ngood and writing ngood
 *                         samples at d_out + b*ngood (ngood = nfft - n_imp_resp + 1); *nprocessed = blocks*ngood. */
#ifndef FIXED_POINT
typedef struct kiss_fastconv_state *kiss_fastconv_cfg;
kiss_fastconv_cfg KISS_FFT_API kiss_fastconv_alloc(const kiss_fft_cpx *imp_resp, size_t n_imp_resp, size_t *pnfft);
void KISS_FFT_API kiss_fastconv_free(kiss_fastconv_cfg cfg);
size_t KISS_FFT_API kiss_fastconv_block_advance(kiss_fastconv_cfg cfg);
size_t KISS_FFT_API kiss_fastconv_nfft(kiss_fastconv_cfg cfg);
int KISS_FFT_API kiss_fastconv_dev(kiss_fastconv_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t n,
                                   size_t *nprocessed, void *stream);
#endif

/* ---- introspection ---------------------------------------------------------------------------------------- */
const char KISS_FFT_API *kiss_fft_cuda_last_error(void);
/* kernels launched by this library since it was loaded */
long long KISS_FFT_API kiss_fft_cuda_launch_count(void);
/* 1 when nfft has a compile-time fused plan (single HBM round trip) in this build, 0 when it runs on the
 * run-time shared-memory kernel, -1 when it is not supported */
int KISS_FFT_API kiss_fft_cuda_plan_kind(int nfft);
/* testing aid: route every length through the run-time shared-memory kernel (1) or restore the default (0) */
void KISS_FFT_API kiss_fft_cuda_force_generic(int on);
/* cap the number of CTAs of the fused kernels launched after this call (0 = fill the device).  Used by the slab
 * transform so that its NVLink-bound peer-store launches share the SMs with HBM-bound launches on another stream. */
void KISS_FFT_API kiss_fft_cuda_set_grid_limit(int max_ctas);
/* sizeof(kiss_fft_scalar) of this build (4 float, 8 double, 2 Q15, 4 Q31) and 1 for fixed point */
int KISS_FFT_API kiss_fft_cuda_scalar_bytes(void);
int KISS_FFT_API kiss_fft_cuda_is_fixed_point(void);

#ifdef __cplusplus
}
#endif
#endif
