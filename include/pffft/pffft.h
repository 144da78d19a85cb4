/*
 * pffft.h -- single-precision C-ABI of the B200-native FFT engine.
 *
 * Drop-in boundary: every prototype below has the same name, argument list,
 * enum values and calling convention as the reference library's
 * include/pffft/pffft.h:124-250 (marton78/pffft).  A program compiled against
 * the reference header links against libpffft_b200.so unchanged; what runs
 * underneath is hand-written sm_100a CUDA, not SSE passes.
 *
 * Pointer rule (the one semantic extension): every float* argument may be a
 * host pointer (the reference's only mode; data is staged over PCIe and the
 * call returns when `output` is complete) or a CUDA device pointer (work is
 * enqueued on the setup's stream, see pffft_b200.h).  Host and device pointers
 * must not be mixed within one call.
 *
 * Data layouts (reference: include/pffft/pffft.h:127-178):
 *   complex, N points : 2N floats, (re,im) interleaved, natural bin order.
 *   real,    N points : N floats of time samples; spectrum = N/2 complex slots,
 *                       slot 0 = (X[0].re, X[N/2].re), slot k = X[k].
 *   "unordered"/z-domain output of pffft_transform: the reference's 4-lane
 *   internal layout, reproduced bit-for-bit so spectra are interchangeable
 *   with the CPU library (pffft_priv_impl.h:1158-1193).
 * Transforms are unnormalised: BACKWARD(FORWARD(x)) == N*x.
 */
#ifndef PFFFT_H
#define PFFFT_H

#include <stddef.h> /* size_t */

#if defined(__GNUC__) || defined(__clang__)
#  define PFFFT_EXPORT __attribute__((visibility("default")))
#else
#  define PFFFT_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque plan: immutable after creation, shareable between host threads. */
typedef struct PFFFT_Setup PFFFT_Setup;

#ifndef PFFFT_COMMON_ENUMS
#define PFFFT_COMMON_ENUMS
/* ABI values fixed by the reference (pffft.h:108-117): FORWARD=0, BACKWARD=1; REAL=0, COMPLEX=1 */
typedef enum { PFFFT_FORWARD, PFFFT_BACKWARD } pffft_direction_t;
typedef enum { PFFFT_REAL, PFFFT_COMPLEX } pffft_transform_t;
#endif

/* ref pffft.h:124 -- NULL when N<0, N>2^26, N%16 (complex) / N%32 (real) != 0,
   or N/4 has a prime factor other than 2,3,5 (pffft_priv_impl.h:1062-1112). */
PFFFT_EXPORT PFFFT_Setup *pffft_new_setup(int N, pffft_transform_t transform);
/* ref pffft.h:125 -- NULL-safe. */
PFFFT_EXPORT void pffft_destroy_setup(PFFFT_Setup *setup);

/* ref pffft.h:157 -- z-domain (unordered) result.  `work` is accepted for ABI
   compatibility and ignored: device scratch is owned by the setup. input==output allowed. */
PFFFT_EXPORT void pffft_transform(PFFFT_Setup *setup, const float *input, float *output,
                                  float *work, pffft_direction_t direction);
/* ref pffft.h:166 -- canonical (ordered) result. input==output allowed. */
PFFFT_EXPORT void pffft_transform_ordered(PFFFT_Setup *setup, const float *input, float *output,
                                          float *work, pffft_direction_t direction);
/* ref pffft.h:180 -- FORWARD: z-domain -> canonical, BACKWARD: canonical -> z-domain.
   input and output must not alias. */
PFFFT_EXPORT void pffft_zreorder(PFFFT_Setup *setup, const float *input, float *output,
                                 pffft_direction_t direction);
/* ref pffft.h:195 -- dft_ab += (dft_a * dft_b) * scaling on z-domain spectra; pointers may alias. */
PFFFT_EXPORT void pffft_zconvolve_accumulate(PFFFT_Setup *setup, const float *dft_a,
                                             const float *dft_b, float *dft_ab, float scaling);
/* ref pffft.h:209 -- dft_ab  = (dft_a * dft_b) * scaling. */
PFFFT_EXPORT void pffft_zconvolve_no_accu(PFFFT_Setup *setup, const float *dft_a,
                                          const float *dft_b, float *dft_ab, float scaling);

/* ref pffft.h:212,215 -- layout granularity stays 4 so size rules match the reference;
   the arch string names the actual backend ("sm_100a"). */
PFFFT_EXPORT int pffft_simd_size(void);
PFFFT_EXPORT const char *pffft_simd_arch(void);

/* size algebra, ref pffft.h:220-241 / pffft_priv_impl.h:78-114, pffft_common.c:47-55 */
PFFFT_EXPORT int pffft_min_fft_size(pffft_transform_t transform);
PFFFT_EXPORT int pffft_next_power_of_two(int N);
PFFFT_EXPORT int pffft_is_power_of_two(int N);
PFFFT_EXPORT int pffft_is_valid_size(int N, pffft_transform_t cplx);
PFFFT_EXPORT int pffft_nearest_transform_size(int N, pffft_transform_t cplx, int higher);

/* ref pffft.h:248-249 -- 64-byte aligned host memory.  Here it is page-locked (pinned)
   when a CUDA device is present, so host-pointer transforms DMA without a bounce copy. */
PFFFT_EXPORT void *pffft_aligned_malloc(size_t nb_bytes);
PFFFT_EXPORT void pffft_aligned_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* PFFFT_H */
