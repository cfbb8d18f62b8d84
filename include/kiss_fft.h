/*
 * kiss_fft.h -- 1-D complex transform API of kissfft-b200 (B200 / sm_100a implementation).
 *
 * Source-compatible with the reference's kiss_fft.h: same scalar-type selection macros, same kiss_fft_cpx
 * layout, same entry points with the same argument meaning and error behaviour, so a program written against
 * the reference recompiles and links against libkissfft-<type>.so from this tree unchanged.
 *
 *   entry point (this file)        reference interface it replaces
 *   ----------------------------   --------------------------------------------
 *   kiss_fft_scalar / kiss_fft_cpx  kiss_fft.h:73-90
 *   kiss_fft_alloc                  kiss_fft.h:117, kiss_fft.c:337-372
 *   kiss_fft                        kiss_fft.h:129, kiss_fft.c:401-404
 *   kiss_fft_stride                 kiss_fft.h:134, kiss_fft.c:375-399
 *   kiss_fft_free                   kiss_fft.h:138 (plain free())
 *   kiss_fft_cleanup                kiss_fft.h:144, kiss_fft.c:407-410
 *   kiss_fft_next_fast_size         kiss_fft.h:150, kiss_fft.c:412-424
 *
 * Pointers handed to the transform calls may be ordinary host pointers (the call stages the data through the
 * GPU and returns when the result is in `fout`, like the reference) or CUDA device pointers (the transform is
 * enqueued on the calling thread's default stream and runs in place on the device).  The batched
 * device-pointer variants live in kiss_fft_cuda.h.
 */
#ifndef KISS_FFT_H
#define KISS_FFT_H

/* the same four standard headers the reference's kiss_fft.h pulls in (kiss_fft.h:12-15): programs written against it
 * rely on getting printf / memset / cos / M_PI through this header */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef KISS_FFT_SHARED
# define KISS_FFT_API __attribute__((visibility("default")))
#else
# define KISS_FFT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Datatype of this build -- must match the library that is linked (libkissfft-float / -double / -int16_t /
 * -int32_t), exactly as with the reference. */
#ifdef FIXED_POINT
# include <stdint.h>
# if (FIXED_POINT == 32)
#  define kiss_fft_scalar int32_t /* Q31 */
# else
#  define kiss_fft_scalar int16_t /* Q15 */
# endif
#else
# ifndef kiss_fft_scalar
#  define kiss_fft_scalar float
# endif
#endif

#ifndef KISS_FFT_MALLOC
# define KISS_FFT_MALLOC malloc
#endif
#ifndef KISS_FFT_FREE
# define KISS_FFT_FREE free
#endif

typedef struct {
    kiss_fft_scalar r;
    kiss_fft_scalar i;
} kiss_fft_cpx;

typedef struct kiss_fft_state *kiss_fft_cfg;

/*
 * Plan an nfft-point forward (inverse_fft == 0) or inverse transform.
 *   lenmem == NULL            : the cfg is malloc()ed; release it with kiss_fft_free() / free().
 *   lenmem != NULL            : *lenmem receives the number of bytes needed; the cfg is placed in `mem` when
 *                               mem != NULL and the incoming *lenmem was large enough, otherwise NULL is returned.
 * The cfg is a single relocatable-free POD block (no device handles inside); GPU-side tables are cached
 * internally per (device, nfft, direction) and released by kiss_fft_cleanup().
 */
kiss_fft_cfg KISS_FFT_API kiss_fft_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);

/* fout[0..nfft) = DFT(fin[0..nfft)); float/double results are unscaled in both directions, the fixed-point
 * builds scale by 1/nfft in both directions.  fin == fout is allowed. */
void KISS_FFT_API kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);

/* same, reading the input from every fin_stride-th element */
void KISS_FFT_API kiss_fft_stride(kiss_fft_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout, int fin_stride);

#define kiss_fft_free KISS_FFT_FREE

/* releases every GPU resource the library holds (twiddle tables, staging buffers, streams) */
void KISS_FFT_API kiss_fft_cleanup(void);

/* smallest k >= n whose only prime factors are 2, 3 and 5 */
int KISS_FFT_API kiss_fft_next_fast_size(int n);

#define kiss_fftr_next_fast_size_real(n) (kiss_fft_next_fast_size(((n) + 1) >> 1) << 1)

#ifdef __cplusplus
}
#endif
#endif
