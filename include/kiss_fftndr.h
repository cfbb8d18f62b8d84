/*
 * kiss_fftndr.h -- N-dimensional real transforms of kissfft-b200.
 *
 *   kiss_fftndr_alloc  replaces reference kiss_fftndr.h:23, kiss_fftndr.c:30-84
 *   kiss_fftndr        replaces reference kiss_fftndr.h:32, kiss_fftndr.c:86-110
 *   kiss_fftndri       replaces reference kiss_fftndr.h:42, kiss_fftndr.c:112-132
 *   kiss_fftndr_free   reference kiss_fftndr.h:49 (plain free())
 *
 * The real axis is the LAST one (dims[ndims-1], must be even) as in the reference's implementation
 * (kiss_fftndr.c:36).  kiss_fftndr maps dims[0] x ... x dims[ndims-1] scalars to
 * dims[0] x ... x (dims[ndims-1]/2+1) complex bins (== numpy.fft.rfftn); kiss_fftndri is the reverse.
 */
#ifndef KISS_NDR_H
#define KISS_NDR_H

#include "kiss_fft.h"
#include "kiss_fftnd.h"
#include "kiss_fftr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kiss_fftndr_state *kiss_fftndr_cfg;

kiss_fftndr_cfg KISS_FFT_API kiss_fftndr_alloc(const int *dims, int ndims, int inverse_fft, void *mem, size_t *lenmem);
void KISS_FFT_API kiss_fftndr(kiss_fftndr_cfg cfg, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata);
void KISS_FFT_API kiss_fftndri(kiss_fftndr_cfg cfg, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata);

#define kiss_fftndr_free free

#ifdef __cplusplus
}
#endif
#endif
