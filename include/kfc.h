/*
 * kfc.h -- cfg cache convenience layer of kissfft-b200.
 *
 *   kfc_fft / kfc_ifft / kfc_cleanup   replace reference kfc.h:36-40, kfc.c:24-83
 *
 * One cached plan per (nfft, direction), created on first use, found by a linear search like the reference's.
 * Unlike the reference (unsynchronised globals, kfc.c:21-22) the cache is mutex protected.  kfc_cleanup() frees the
 * cached cfgs; the GPU-side tables they refer to are released by kiss_fft_cleanup().
 */
#ifndef KFC_H
#define KFC_H
#include "kiss_fft.h"

#ifdef __cplusplus
extern "C" {
#endif

void KISS_FFT_API kfc_fft(int nfft, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);
void KISS_FFT_API kfc_ifft(int nfft, const kiss_fft_cpx *fin, kiss_fft_cpx *fout);
void KISS_FFT_API kfc_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
