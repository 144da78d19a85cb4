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

/* ---- streaming and partitioned convolution (callers of pffastconv_apply / pffft_zconvolve_accumulate) ---- */
/* Stateful form of pffastconv_apply's contract ("returns the number of produced samples; feed the rest again", ref
   include/pffft/pffastconv.h:160-171): the setup keeps the unconsumed tail of the stream ON THE DEVICE between calls.
   push: appends cplxInputLen (complex) samples, writes the outputs of every block that became complete and returns their
   count (<= outputCapacity, else an error); flush: outputs for everything fed so far, including the partial last block
   (applyFlush = 1); the last filterLen-1 samples stay pending so the stream may continue.  Any chunking followed by one
   flush produces bit-identical samples to ONE pffastconv_apply(flush=1) over the whole stream.  Host or device pointers;
   returns -1 on error. */
PFFFT_EXPORT int pffastconvb_push(PFFASTCONV_Setup *setup, const float *input, int cplxInputLen, float *output, int outputCapacity);
PFFFT_EXPORT int pffastconvb_flush(PFFASTCONV_Setup *setup, float *output, int outputCapacity);
PFFFT_EXPORT int pffastconvb_pending(const PFFASTCONV_Setup *setup);   /* samples waiting for more input */
PFFFT_EXPORT void pffastconvb_reset(PFFASTCONV_Setup *setup);          /* forget the pending samples (new stream) */

/* Uniformly partitioned overlap-save convolution for LONG real filters with low latency: taps in partitions of partLen (a
   power of two >= 16), transforms of 2*partLen points, pffft_zconvolve_accumulate as the inner loop (ref pffft.h:182-195)
   fused over the partitions.  Output convention of pffastconv: y[n] = sum_j x[n+j]*taps[filterLen-1-j], n in [0, len-filterLen].
   apply returns the number of outputs written (0 when len < filterLen), < 0 on error; host or device pointers. */
typedef struct PFFASTCONVB_Partitioned PFFASTCONVB_Partitioned;
PFFFT_EXPORT PFFASTCONVB_Partitioned *pffastconvb_partitioned_new(const float *taps, int filterLen, int partLen);
PFFFT_EXPORT void pffastconvb_partitioned_destroy(PFFASTCONVB_Partitioned *c);
PFFFT_EXPORT long long pffastconvb_partitioned_apply(PFFASTCONVB_Partitioned *c, const float *input, long long len, float *output);
PFFFT_EXPORT int pffastconvb_partitioned_partitions(const PFFASTCONVB_Partitioned *c);
PFFFT_EXPORT int pffastconvb_partitioned_set_stream(PFFASTCONVB_Partitioned *c, void *cuda_stream);

/* ---- multi-GPU: the batch is sharded over the GPUs of one node, the plan tables are broadcast ONCE over
   NCCL / NVLink, and nothing else crosses GPUs (every transform is independent, ref pffft.h:102-106) ---- */
/* Device address and size of the plan's twiddle/rotation tables (the analogue of the reference's
   PFFFT_Setup::data, pffft_priv_impl.h:1085-1103). */
PFFFT_EXPORT int pffftb_setup_tables(PFFFT_Setup *setup, void **device_ptr, size_t *nbytes);
PFFFT_EXPORT int pffftdb_setup_tables(PFFFTD_Setup *setup, void **device_ptr, size_t *nbytes);

/* (a) ONE PROCESS, ALL GPUS.  pffftb_multi_new builds one plan per GPU (ngpus <= 0: every visible device), creates the
   communicators with ncclCommInitAll and overwrites the tables of GPUs 1.. with GPU 0's by one ncclBroadcast (root 0), each
   GPU on its own stream, so all GPUs use bit-identical tables.  NCCL is loaded at run time (libnccl.so.2); when it is
   absent the same broadcast is done with cudaMemcpyPeer and pffftb_multi_broadcast_backend() says so. */
typedef struct PFFFTB_Multi PFFFTB_Multi;
PFFFT_EXPORT PFFFTB_Multi *pffftb_multi_new(int N, pffft_transform_t transform, int ngpus);
PFFFT_EXPORT void pffftb_multi_destroy(PFFFTB_Multi *m);
PFFFT_EXPORT int pffftb_multi_ngpus(const PFFFTB_Multi *m);
PFFFT_EXPORT PFFFT_Setup *pffftb_multi_setup(PFFFTB_Multi *m, int gpu);        /* the plan living on GPU `gpu` */
PFFFT_EXPORT const char *pffftb_multi_broadcast_backend(const PFFFTB_Multi *m); /* "nccl" or "memcpy_peer" */
/* HOST pointers: contiguous ranges [g*batch/G, (g+1)*batch/G) go to GPU g, each through that plan's three-stream pipeline
   driven by its own host thread; returns when `output` is complete. */
PFFFT_EXPORT int pffftb_multi_transform_batch(PFFFTB_Multi *m, const float *input, float *output, size_t batch,
                                              pffft_direction_t direction, int ordered);
/* DEVICE-resident shards: input[g]/output[g] are device pointers on GPU g holding batch[g] transforms; enqueued on every
   GPU's plan stream, returns at once; pffftb_multi_synchronize waits for all of them. */
PFFFT_EXPORT int pffftb_multi_transform_shards(PFFFTB_Multi *m, const float *const *input, float *const *output,
                                               const size_t *batch, pffft_direction_t direction, int ordered);
PFFFT_EXPORT int pffftb_multi_synchronize(PFFFTB_Multi *m);

/* (b) ONE PROCESS PER GPU (torchrun, MPI): rank 0 calls pffftb_nccl_unique_id and ships the 128 bytes to the other ranks
   by any means; every rank then calls pffftb_setup_broadcast_tables with the same id -- ncclCommInitRank + one
   ncclBroadcast(root 0) of the tables inside the library. */
PFFFT_EXPORT int pffftb_nccl_unique_id(void *id128);
PFFFT_EXPORT int pffftb_setup_broadcast_tables(PFFFT_Setup *setup, const void *id128, int rank, int nranks);

/* ---- diagnostics ---- */
PFFFT_EXPORT const char *pffftb_last_error(void);            /* thread-local, "" when none */
PFFFT_EXPORT unsigned long long pffftb_launch_count(void);   /* kernels launched by this library so far */
PFFFT_EXPORT int pffftb_device_synchronize(void);            /* cudaDeviceSynchronize passthrough for C callers */

#ifdef __cplusplus
}
#endif
#endif /* PFFFT_B200_H */
