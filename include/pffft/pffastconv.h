/*
 * pffastconv.h -- overlap-save FIR convolution on top of the FFT engine.
 *
 * Same symbols, flag values and semantics as the reference's
 * include/pffft/pffastconv.h:83-180 / src/pffastconv.c:58-263:
 *   y[n] = sum_{j<filterLen} x[n+j] * h[filterLen-1-j],  n = 0 .. inputLen-filterLen
 * (no flip with PFFASTCONV_CORRELATION).  Here every block of the stream is an
 * independent unit of work and all blocks of one pffastconv_apply call are
 * processed by one batched GPU pass (load -> real FFT -> x Hf/Nfft -> inverse
 * -> store valid samples) instead of the reference's sequential per-block loop.
 * `input`/`output` may be host or device pointers (see pffft.h).
 */
#ifndef PFFASTCONV_H
#define PFFASTCONV_H

#include <stddef.h>
#include "pffft.h"

#if defined(__GNUC__) || defined(__clang__)
#  define PFFASTCONV_EXPORT __attribute__((visibility("default")))
#else
#  define PFFASTCONV_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Holds the filter spectrum and device scratch: one setup per filter and per calling thread
   (ref pffastconv.h:77-81). */
typedef struct PFFASTCONV_Setup PFFASTCONV_Setup;

/* bit values fixed by the reference (pffastconv.h:83-134) */
typedef enum {
  PFFASTCONV_CPLX_INP_OUT    = 1,  /* input/output are interleaved complex streams (inputLen complex samples) */
  PFFASTCONV_CPLX_FILTER     = 2,  /* complex taps: unsupported, new_setup returns NULL (ref pffastconv.c:71-72) */
  PFFASTCONV_DIRECT_INP      = 4,  /* copy-elision hint of the CPU library; results identical, ignored here */
  PFFASTCONV_DIRECT_OUT      = 8,  /* copy-elision hint of the CPU library; results identical, ignored here */
  PFFASTCONV_CPLX_SINGLE_FFT = 16, /* with CPLX_INP_OUT: one real FFT of twice the length over the interleaved stream */
  PFFASTCONV_SYMMETRIC       = 32, /* informational */
  PFFASTCONV_CORRELATION     = 64  /* taps are used as given (no time reversal) */
} pffastconv_flags_t;

/* ref pffastconv.h:145 / pffastconv.c:58-116.  *blockLen is in-out: on return it holds the
   FFT block length actually used = max(2*nextpow2(filterLen-1), 32, nextpow2(*blockLen)). */
PFFASTCONV_EXPORT PFFASTCONV_Setup *pffastconv_new_setup(const float *filterCoeffs, int filterLen,
                                                         int *blockLen, int flags);
/* ref pffastconv.h:147 -- NULL-safe. */
PFFASTCONV_EXPORT void pffastconv_destroy_setup(PFFASTCONV_Setup *setup);

/* ref pffastconv.h:173 / pffastconv.c:133-263.  Returns the number of (complex) output samples
   written, which is also the number of input samples consumed; with applyFlush==0 only whole
   blocks are processed and the caller re-feeds the unconsumed tail. */
PFFASTCONV_EXPORT int pffastconv_apply(PFFASTCONV_Setup *setup, const float *input, int inputLen,
                                       float *output, int applyFlush);

/* ref pffastconv.h:175-176 */
PFFASTCONV_EXPORT void *pffastconv_malloc(size_t nb_bytes);
PFFASTCONV_EXPORT void pffastconv_free(void *ptr);
/* ref pffastconv.h:179 */
PFFASTCONV_EXPORT int pffastconv_simd_size(void);

#ifdef __cplusplus
}
#endif
#endif /* PFFASTCONV_H */
