/* multi_gpu_c2c.c -- every GPU of the node from ONE plain C process (include/pffft/pffft_b200.h, multi-GPU section).
 *
 *   cc -Iinclude/pffft examples/multi_gpu_c2c.c -Lpffft_b200 -lpffft_b200 -Wl,-rpath,$PWD/pffft_b200 -lm -o multi_gpu_c2c
 *   ./multi_gpu_c2c [ngpus] [log2_batch]
 *
 * pffftb_multi_new builds one plan per GPU and broadcasts GPU 0's twiddle tables to the others with one ncclBroadcast
 * (ncclCommInitAll, one stream per device).  The batch of N = 1024 complex transforms is sharded in contiguous ranges:
 * (1) from page-locked HOST memory through every GPU's copy/compute pipeline, (2) from device-resident shards.
 * Nothing else crosses GPUs: transforms are independent (reference: include/pffft/pffft.h:102-106).                  */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "pffft_b200.h"

int main(int argc, char **argv) {
  const int N = 1024;
  const int want = argc > 1 ? atoi(argv[1]) : 0;
  const size_t batch = (size_t)1 << (argc > 2 ? atoi(argv[2]) : 14), per = 2 * (size_t)N;
  PFFFTB_Multi *m = pffftb_multi_new(N, PFFFT_COMPLEX, want);
  if (!m) { fprintf(stderr, "pffftb_multi_new failed: %s\n", pffftb_last_error()); return 1; }
  const int G = pffftb_multi_ngpus(m);
  float *x = (float *)pffft_aligned_malloc(batch * per * sizeof(float));
  float *y = (float *)pffft_aligned_malloc(batch * per * sizeof(float));
  float *z = (float *)pffft_aligned_malloc(batch * per * sizeof(float));
  if (!x || !y || !z) return 2;
  for (size_t i = 0; i < batch * per; ++i) x[i] = (float)rand() / (float)RAND_MAX * 2.f - 1.f;
  if (pffftb_multi_transform_batch(m, x, y, batch, PFFFT_FORWARD, 1) ||
      pffftb_multi_transform_batch(m, y, z, batch, PFFFT_BACKWARD, 1)) {
    fprintf(stderr, "multi transform failed: %s\n", pffftb_last_error());
    return 3;
  }
  double worst = 0.0;
  for (size_t i = 0; i < batch * per; ++i) { const double e = fabs((double)z[i] / N - (double)x[i]); if (e > worst) worst = e; }
  printf("%d GPU(s), tables broadcast by %s: %zu x N=%d complex from host memory, max |ifft(fft(x))/N - x| = %.3g\n",
         G, pffftb_multi_broadcast_backend(m), batch, N, worst);
  /* every GPU must give bit-identical results for the same input (same tables, same kernels) */
  int same = 1;
  {
    float *a = (float *)pffft_aligned_malloc(per * sizeof(float)), *b = (float *)pffft_aligned_malloc(per * sizeof(float));
    pffft_transform_ordered(pffftb_multi_setup(m, 0), x, a, NULL, PFFFT_FORWARD);
    for (int g = 1; g < G; ++g) {
      pffft_transform_ordered(pffftb_multi_setup(m, g), x, b, NULL, PFFFT_FORWARD);
      for (size_t i = 0; i < per; ++i) if (a[i] != b[i]) same = 0;
    }
    pffft_aligned_free(a); pffft_aligned_free(b);
  }
  printf("results of all GPUs bit-identical: %s\n", same ? "yes" : "NO");
  pffft_aligned_free(x); pffft_aligned_free(y); pffft_aligned_free(z);
  pffftb_multi_destroy(m);
  return (worst < 1e-5 && same) ? 0 : 4;
}
