/*
 * kiss_fftr.h -- real-input / real-output 1-D transforms of kissfft-b200.
 *
 *   kiss_fftr_alloc   replaces reference kiss_fftr.h:29,  kiss_fftr.c:21-61
 *   kiss_fftr         replaces reference kiss_fftr.h:37,  kiss_fftr.c:63-117
 *   kiss_fftri        replaces reference kiss_fftr.h:43,  kiss_fftr.c:119-155
 *   kiss_fftr_free    reference kiss_fftr.h:49 (plain free())
 *
 * nfft must be even.  kiss_fftr reads nfft scalars and writes nfft/2+1 complex bins; kiss_fftri does the
 * opposite.  A cfg planned with inverse_fft == 0 only serves kiss_fftr and one planned with inverse_fft != 0
 * only serves kiss_fftri; the wrong pairing is a logged no-op exactly as in the reference.
 */
#ifndef KISS_FFTR_H
#define KISS_FFTR_H

#include "kiss_fft.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kiss_fftr_state *kiss_fftr_cfg;

kiss_fftr_cfg KISS_FFT_API kiss_fftr_alloc(int nfft, int inverse_fft, void *mem, size_t *lenmem);
void KISS_FFT_API kiss_fftr(kiss_fftr_cfg cfg, const kiss_fft_scalar *timedata, kiss_fft_cpx *freqdata);
void KISS_FFT_API kiss_fftri(kiss_fftr_cfg cfg, const kiss_fft_cpx *freqdata, kiss_fft_scalar *timedata);

#define kiss_fftr_free KISS_FFT_FREE

#ifdef __cplusplus
}
#endif
#endif
