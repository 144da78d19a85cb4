/*
 * pffft_b200.h -- what the reference C-ABI cannot express: batches, device
 * residency, streams, multi-GPU table sharing and error reporting.
 *
 * The reference (marton78/pffft) transforms ONE vector per call from host
 * memory (include/pffft/pffft.h:157,166) and returns void.  On a B200 the unit
 * of work is a batch of independent transforms resident in HBM; these entry
 * points add exactly that and nothing else.  They are plain C (pointers and
 * sizes), so the same cgo/ctypes/JNI style binding used for pffft.h applies.
 *
 * Batch layout: batch-major contiguous, transform b starts at element
 * b*pffftb_floats_per_transform(setup) (N floats real, 2N floats complex); every
 * transform individually uses the layouts documented in pffft.h.
 *
 * All functions returning int return 0 on success, a non-zero CUDA error code
 * otherwise (text via pffftb_last_error()).  Nothing here falls back to the CPU:
 * without a usable sm_100 device, setup creation fails.
 */
#ifndef PFFFT_B200_H
#define PFFFT_B200_H

#include <stddef.h>
#include "pffft.h"
#include "pffft_double.h"
#include "pffastconv.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- batched transforms (extends pffft_transform / pffft_transform_ordered, ref pffft.h:157,166) ---- */
/* `input`/`output`: both host or both device pointers, batch*floats_per_transform elements each;
   input==output allowed.  Host pointers: the batch is streamed through the GPU in chunks with
   H2D copy, kernels and D2H copy overlapped on three streams; the call returns when `output`
   is complete.  Device pointers: kernels are enqueued on the setup's stream and the call returns
   immediately.  ordered!=0 -> canonical layout, ordered==0 -> z-domain layout. */
PFFFT_EXPORT int pffftb_transform_batch(PFFFT_Setup *setup, const float *input, float *output,
                                        size_t batch, pffft_direction_t direction, int ordered);
PFFFT_EXPORT int pffftdb_transform_batch(PFFFTD_Setup *setup, const double *input, double *output,
                                         size_t batch, pffft_direction_t direction, int ordered);

/* batched pffft_zreorder (ref pffft.h:180); input must not alias output. */
PFFFT_EXPORT int pffftb_zreorder_batch(PFFFT_Setup *setup, const float *input, float *output,
                                       size_t batch, pffft_direction_t direction);
PFFFT_EXPORT int pffftdb_zreorder_batch(PFFFTD_Setup *setup, const double *input, double *output,
                                        size_t batch, pffft_direction_t direction);

/* batched pffft_zconvolve_{accumulate,no_accu} (ref pffft.h:195,209).  dft_a and dft_ab hold
   `batch` spectra; dft_b holds `batch` spectra, or ONE spectrum applied to every element of the
   batch when b_is_shared!=0 (the filter case).  accumulate!=0: ab += a*b*scaling, else ab = a*b*scaling. */
PFFFT_EXPORT int pffftb_zconvolve_batch(PFFFT_Setup *setup, const float *dft_a, const float *dft_b,
                                        float *dft_ab, float scaling, size_t batch, int b_is_shared,
                                        int accumulate);
PFFFT_EXPORT int pffftdb_zconvolve_batch(PFFFTD_Setup *setup, const double *dft_a, const double *dft_b,
                                         double *dft_ab, double scaling, size_t batch, int b_is_shared,
                                         int accumulate);

/* ---- plan introspection ---- */
PFFFT_EXPORT size_t pffftb_floats_per_transform(const PFFFT_Setup *setup);   /* N (real) or 2N (complex) */
PFFFT_EXPORT size_t pffftdb_doubles_per_transform(const PFFFTD_Setup *setup);
PFFFT_EXPORT int pffftb_setup_device(const PFFFT_Setup *setup);              /* CUDA ordinal the plan lives on */
PFFFT_EXPORT const char *pffftb_setup_kernel(const PFFFT_Setup *setup);      /* name of the kernel family chosen */
PFFFT_EXPORT const char *pffftdb_setup_kernel(const PFFFTD_Setup *setup);

/* ---- streams ---- */
/* cudaStream_t (as void*) used for device-pointer calls; default is the legacy default stream (0),
   which is also PyTorch's default stream. */
PFFFT_EXPORT int pffftb_set_stream(PFFFT_Setup *setup, void *cuda_stream);
PFFFT_EXPORT int pffftdb_set_stream(PFFFTD_Setup *setup, void *cuda_stream);
PFFFT_EXPORT int pffastconvb_set_stream(PFFASTCONV_Setup *setup, void *cuda_stream);

/* ---- multi-GPU: one process per GPU, batch sharded, tables broadcast once ---- */
/* Device address and size of the plan's twiddle/rotation tables (the analogue of the reference's
   PFFFT_Setup::data, pffft_priv_impl.h:1085-1103).  Rank 0 builds them; the other ranks overwrite
   theirs with one ncclBroadcast(root 0) over NVLink so every GPU uses bit-identical tables.
   No other inter-GPU traffic exists on this path. */
PFFFT_EXPORT int pffftb_setup_tables(PFFFT_Setup *setup, void **device_ptr, size_t *nbytes);
PFFFT_EXPORT int pffftdb_setup_tables(PFFFTD_Setup *setup, void **device_ptr, size_t *nbytes);

/* ---- diagnostics ---- */
PFFFT_EXPORT const char *pffftb_last_error(void);            /* thread-local, "" when none */
PFFFT_EXPORT unsigned long long pffftb_launch_count(void);   /* kernels launched by this library so far */
PFFFT_EXPORT int pffftb_device_synchronize(void);            /* cudaDeviceSynchronize passthrough for C callers */

#ifdef __cplusplus
}
#endif
#endif /* PFFFT_B200_H */
