/*
 * kiss_fft_cuda.h -- batched / device-pointer extension of the kissfft API (new in kissfft-b200).
 *
 * The reference has no batched or device interface; its callers loop kiss_fft()/kiss_fftr() over rows
 * (test/benchkiss.c:83-115, tools/fftutil.c:19-126).  These entry points are that loop, executed by one kernel
 * launch: they take the same cfg objects the reference-compatible allocators return.
 *
 * Conventions
 *   - `*_dev` functions take CUDA device pointers and a cudaStream_t (passed as void*, NULL = default stream);
 *     they enqueue work and return without synchronising -- EXCEPT where they need internal scratch memory that goes
 *     back to a shared pool: rows longer than the shared-memory kernels hold (nfft > ~14.5 k complex float),
 *     kiss_fftnd_dev with d_work == NULL when an axis has no layout-keeping plan or when a large 3-D float / double
 *     array takes the three plane-local transposing passes (pass d_work to keep that case stream-ordered),
 *     kiss_fftndri_dev, kiss_fftndr_dev in the fallback case, and the unfused fast convolution.  Those wait for the
 *     stream before returning.
 *   - functions without the suffix take HOST pointers and return when the result is in the output buffer.  The batch is
 *     cut into chunks that move through parallel lanes (H2D, kernel, D2H on one stream per lane).  Pinned (or
 *     cudaHostRegister-ed) caller buffers are handed to the copy engines directly; ordinary pageable buffers -- what
 *     callers of the reference pass -- are bounced through internal pinned buffers by the lanes' host threads, so both
 *     copy directions and the kernels still overlap (a cudaMemcpyAsync on pageable memory would not).
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

/* the same with the peers' column blocks spread out in the input: block s reads the input columns
 * s*peer_col_dist + [0, cols_per_peer) of every plane (peer_col_dist >= cols_per_peer).  Lets one launch handle a CHUNK
 * of every destination's column range, so that the exchange pipelines chunk by chunk (kiss_fftnd_mgpu_exec).
 * out_col_dist > 0: row (plane p, column c of block s) lands at d_peers[s][p*out_plane_dist + c*out_col_dist + k] (0: the
 * rows of a plane back to back, c*nfft) -- lets the receiver keep all planes of one column together.
 * max_ctas > 0 caps the persistent grid of this launch: a pass whose stores cross NVLink is link-bound and needs only a
 * fraction of the SMs, the rest stay free for HBM-bound kernels on other streams. */
int KISS_FFT_API kiss_fft_planes_pass_peers2_dev(kiss_fft_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *const *d_peers,
                                                 int npeers, size_t nplanes, size_t cols_per_peer, size_t peer_col_dist,
                                                 size_t col_stride, size_t in_plane_dist, size_t out_plane_dist, size_t out_col_dist,
                                                 int max_ctas, void *stream);

/* kiss_fftndr / kiss_fftndri on device buffers (kiss_fftndr.c:86-132) */
int KISS_FFT_API kiss_fftndr_dev(kiss_fftndr_cfg cfg, const kiss_fft_scalar *d_time, kiss_fft_cpx *d_freq, void *stream);
int KISS_FFT_API kiss_fftndri_dev(kiss_fftndr_cfg cfg, const kiss_fft_cpx *d_freq, kiss_fft_scalar *d_time, void *stream);

/* ---- kiss_fftnd over several GPUs: slab decomposition, ONE PROCESS PER GPU ---------------------------
 * The reference's kiss_fftnd (kiss_fftnd.c:156-188) transforms one array in one address space.  Here a 3-D array
 * d0 x d1 x d2 is distributed in slabs of d0/G planes: rank r holds x[r*d0/G .. (r+1)*d0/G)[d1][d2].  One exchange
 * (all-to-all) is needed; the result comes out "transposed": rank r holds X[k0][k1][k2] for k2 in [r*d2/G, (r+1)*d2/G)
 * stored as d_out[k2 - r*d2/G][k1][k0].  dims[0] and dims[2] must be divisible by G (G <= 16).
 * ndims == 2 (d0 x d1, both divisible by G): rank r holds rows r*d0/G ..., receives X[k0][k1] for its k1 range stored
 * d_out[k1 - r*d1/G][k0]; rows in place, one transposing exchange (peer stores or NCCL), rows.
 *
 *   kiss_fftnd_mgpu_get_id   rank 0 only: fills a KISS_FFT_MGPU_ID_BYTES rendezvous id (an ncclUniqueId); the caller
 *                            distributes it to the other ranks by whatever it has (MPI_Bcast, a pipe, a file, ...)
 *   kiss_fftnd_mgpu_alloc    COLLECTIVE; current CUDA device = this rank's GPU.  flags: KISS_FFT_MGPU_P2P asks for the
 *                            fused exchange -- the column-pass kernel stores its rows straight into the other ranks'
 *                            receive buffers over NVLink (buffers mapped with CUDA IPC: all ranks on one node) -- and
 *                            falls back to NCCL (grouped ncclSend/ncclRecv per chunk, libnccl.so.2 loaded at run time)
 *                            when peer mapping is not possible; kiss_fftnd_mgpu_uses_p2p() tells which one is active.
 *                            nranks == 1 needs neither NCCL nor an id.
 *   kiss_fftnd_mgpu_exec     COLLECTIVE, stream-ordered on `stream`: d_in is overwritten (its rows are transformed in
 *                            place, like kiss_fft with fin == fout); d_out receives the transposed-out slab.
 *   kiss_fftnd_mgpu_free     COLLECTIVE. */
/*                            KISS_FFT_MGPU_REFERENCE_ORDER: the exact mode for the fixed-point builds.  kiss_fftnd sweeps the
 *                            axes in the order 0, 1, 2 and Q15 / Q31 results depend on it; this mode keeps that order with
 *                            the same single exchange by starting from slabs along the LAST axis: rank r passes
 *                            x[i0][i1][r*d2/G + c] stored [d0][d1][d2/G] and receives X[r*d0/G + p][k1][k2] stored
 *                            [d0/G][d1][d2] -- natural order, bit-identical to the rows kiss_fftnd writes (every datatype).
 *                            Can be combined with KISS_FFT_MGPU_P2P. */
#define KISS_FFT_MGPU_ID_BYTES 128
#define KISS_FFT_MGPU_P2P 1u
#define KISS_FFT_MGPU_REFERENCE_ORDER 2u
typedef struct kiss_fftnd_mgpu_state *kiss_fftnd_mgpu_cfg;
int KISS_FFT_API kiss_fftnd_mgpu_get_id(void *id);
kiss_fftnd_mgpu_cfg KISS_FFT_API kiss_fftnd_mgpu_alloc(const int *dims, int ndims, int inverse_fft, int rank, int nranks,
                                                       const void *id, unsigned flags);
int KISS_FFT_API kiss_fftnd_mgpu_exec(kiss_fftnd_mgpu_cfg cfg, kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, void *stream);
void KISS_FFT_API kiss_fftnd_mgpu_free(kiss_fftnd_mgpu_cfg cfg);
/* complex elements of this rank's input / output slab; 1 when the peer-store exchange is active; chunks the exchange is
 * pipelined in; bytes this rank sends to other ranks per exec (NVLink roofline numerator); last error of this thread */
size_t KISS_FFT_API kiss_fftnd_mgpu_local_in_elems(kiss_fftnd_mgpu_cfg cfg);
size_t KISS_FFT_API kiss_fftnd_mgpu_local_out_elems(kiss_fftnd_mgpu_cfg cfg);
int KISS_FFT_API kiss_fftnd_mgpu_uses_p2p(kiss_fftnd_mgpu_cfg cfg);
int KISS_FFT_API kiss_fftnd_mgpu_chunks(kiss_fftnd_mgpu_cfg cfg);
size_t KISS_FFT_API kiss_fftnd_mgpu_a2a_bytes(kiss_fftnd_mgpu_cfg cfg);
const char KISS_FFT_API *kiss_fftnd_mgpu_last_error(void);
/* tuning aid, COLLECTIVE with no exec in flight: knobs = {chunks of the exchange, plane groups, CTA cap of the link-bound
 * launches (0 = none), priority stream for them (0/1), SMs the HBM-bound launches leave to them, trace (0/1), SMs of the
 * link-bound green-context partition (0 = no partition), unused, share of the last chunk in sixteenths (0 = equal chunks)},
 * a negative entry keeps the current value; _knob reads one back
 * (6: SMs of the link partition actually provisioned, 7: SMs of the other one) */
int KISS_FFT_API kiss_fftnd_mgpu_tune(kiss_fftnd_mgpu_cfg cfg, const int *knobs, int nknobs);
int KISS_FFT_API kiss_fftnd_mgpu_knob(kiss_fftnd_mgpu_cfg cfg, int which);
/* tuning aid: with knob 5 (trace) set, every launch of an exec is bracketed by timed events; after the exec this writes one
 * "name start_ms end_ms" line per launch (A<group> rows, B<group>.<chunk> column pass + exchange stores, X NCCL exchange,
 * C0.<chunk> axis-0 pass) into buf.  Synchronises the device. */
int KISS_FFT_API kiss_fftnd_mgpu_trace(kiss_fftnd_mgpu_cfg cfg, char *buf, size_t len);
/* testing aid (host logic, no CUDA call): boundaries of the k2-column chunks for `cols` columns per rank */
int KISS_FFT_API kiss_fftnd_mgpu_debug_chunks(int cols, int nranks, int want, int tail16, int *coff, int cap);

/* ---- host-pointer batched (parallel lanes of H2D / kernel / D2H, pinned bounce buffers for pageable memory) --- */
int KISS_FFT_API kiss_fft_batch(kiss_fft_cfg cfg, const kiss_fft_cpx *in, kiss_fft_cpx *out, size_t howmany);
int KISS_FFT_API kiss_fftr_batch(kiss_fftr_cfg cfg, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata, size_t howmany);
int KISS_FFT_API kiss_fftri_batch(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata, size_t howmany);

/* ---- fast convolution (float / double builds only, like the reference: kiss_fastfir.c:152) ---------------------------
 * The overlap-scrap FIR filter of the reference's tools/kiss_fastfir.c: FFT -> multiply by the filter's frequency response
 * -> IFFT for every block.
 *   kiss_fastconv_alloc   kiss_fastfir_alloc (kiss_fastfir.c:59-165), complex samples: *pnfft == 0 picks the reference's
 *                         default size; any nfft the transforms accept
 *   kiss_fastconvr_alloc  the same for the reference's REAL_FASTFIR build (real samples and impulse response,
 *                         kiss_fftr / kiss_fftri, nfft/2+1 bins; nfft even)
 *   kiss_fastconv_dev /   kff_nocopy / fastconv1buf (kiss_fastfir.c:167-206) on device buffers: processes every complete
 *   kiss_fastconvr_dev    nfft-sample block of the n input samples, block b reading d_in + b*ngood and writing ngood
 *                         samples at d_out + b*ngood (ngood = nfft - n_imp_resp + 1); *nprocessed = blocks*ngood.
 * Complex samples with nfft in {256, 512, 1024, 2048, 4096} run as ONE fused kernel (the spectrum never leaves the SM),
 * stream-ordered.  Every other case -- other lengths, real samples -- is composed from separate launches (block gather,
 * batched transform, pointwise product, batched inverse, clipped copy) over internal scratch and waits for the stream. */
#ifndef FIXED_POINT
typedef struct kiss_fastconv_state *kiss_fastconv_cfg;
kiss_fastconv_cfg KISS_FFT_API kiss_fastconv_alloc(const kiss_fft_cpx *imp_resp, size_t n_imp_resp, size_t *pnfft);
kiss_fastconv_cfg KISS_FFT_API kiss_fastconvr_alloc(const kiss_fft_scalar *imp_resp, size_t n_imp_resp, size_t *pnfft);
void KISS_FFT_API kiss_fastconv_free(kiss_fastconv_cfg cfg);
size_t KISS_FFT_API kiss_fastconv_block_advance(kiss_fastconv_cfg cfg);
size_t KISS_FFT_API kiss_fastconv_nfft(kiss_fastconv_cfg cfg);
int KISS_FFT_API kiss_fastconv_dev(kiss_fastconv_cfg cfg, const kiss_fft_cpx *d_in, kiss_fft_cpx *d_out, size_t n,
                                   size_t *nprocessed, void *stream);
int KISS_FFT_API kiss_fastconvr_dev(kiss_fastconv_cfg cfg, const kiss_fft_scalar *d_in, kiss_fft_scalar *d_out, size_t n,
                                    size_t *nprocessed, void *stream);
#endif

/* ---- introspection ---------------------------------------------------------------------------------------- */
const char KISS_FFT_API *kiss_fft_cuda_last_error(void);
/* kernels launched by this library since it was loaded */
long long KISS_FFT_API kiss_fft_cuda_launch_count(void);
/* 1 when nfft has a compile-time fused plan (single HBM round trip) in this build, 0 when it runs on the
 * run-time shared-memory kernel, 2 when it is longer than those hold and takes the long-row paths (four-step or one
 * radix stage per launch), -1 for nfft <= 0 */
int KISS_FFT_API kiss_fft_cuda_plan_kind(int nfft);
/* testing aid: route every length through the run-time shared-memory kernel (1) or restore the default (0) */
void KISS_FFT_API kiss_fft_cuda_force_generic(int on);
/* cap the number of CTAs of the fused kernels the CALLING THREAD launches after this call (0 = fill the device).  Used by the slab
 * transform so that its NVLink-bound peer-store launches share the SMs with HBM-bound launches on another stream. */
void KISS_FFT_API kiss_fft_cuda_set_grid_limit(int max_ctas);
/* testing aid (host logic, no CUDA call): the chunks a host-pointer batched call of `howmany` rows is cut into for a
 * chunk size of `rows` rows -- small chunks at both ends when ramp != 0; returns their number, fills first[] / n[] up to cap */
size_t KISS_FFT_API kiss_fft_cuda_debug_chunks(size_t howmany, size_t rows, int ramp, size_t *first, size_t *n, size_t cap);
/* sizeof(kiss_fft_scalar) of this build (4 float, 8 double, 2 Q15, 4 Q31) and 1 for fixed point */
int KISS_FFT_API kiss_fft_cuda_scalar_bytes(void);
int KISS_FFT_API kiss_fft_cuda_is_fixed_point(void);

#ifdef __cplusplus
}
#endif
#endif
