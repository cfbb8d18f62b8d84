/*
 * kiss_fftnd.h -- N-dimensional complex transform of kissfft-b200.
 *
 *   kiss_fftnd_alloc  replaces reference kiss_fftnd.h:20, kiss_fftnd.c:20-92
 *   kiss_fftnd        replaces reference kiss_fftnd.h:21, kiss_fftnd.c:156-188
 *
 * Row-major dims[0] x dims[1] x ... ; output in natural order (== numpy.fft.fftn).  Axes are processed in the
 * reference's order 0,1,...,ndims-1 (each pass transforms the columns of a dims[k] x rest view and stores them
 * as rows), which is what makes the Q15/Q31 builds bit-identical to the reference.  fin == fout is allowed.
 */
#ifndef KISS_FFTND_H
#define KISS_FFTND_H

#include "kiss_fft.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kiss_fftnd_state *kiss_fftnd_cfg;

kiss_fftnd_cfg KISS_FFT_API kiss_fftnd_alloc(const int *dims, int ndims, int inverse_fft, void *mem, size_t *lenmem);
void KISS_FFT_API kiss_fftnd(kiss_fftnd_cfg cfg, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);

#ifdef __cplusplus
}
#endif
#endif
