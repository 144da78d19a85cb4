/* batch_c2c.c -- a plain C program on the batched extension of the C-ABI (include/pffft/pffft_b200.h).
 *
 *   cc -Iinclude/pffft examples/batch_c2c.c -Lpffft_b200 -lpffft_b200 -Wl,-rpath,$PWD/pffft_b200 -lm -o batch_c2c
 *
 * 4096 forward transforms of N = 1024 complex points from page-locked host memory, then back; prints the round-trip
 * error and which kernel family the plan chose.  The same calls accept device pointers (cudaMalloc) unchanged.
 * The classic single-vector calls (pffft_transform_ordered, ...) are the reference's and need no example: the
 * reference's own examples/ compile against this library unmodified (tests/test_reference_suite_gpu.py).            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "pffft_b200.h"

int main(void) {
  const int N = 1024;
  const size_t batch = 4096, per = 2 * (size_t)N;
  PFFFT_Setup *s = pffft_new_setup(N, PFFFT_COMPLEX);
  if (!s) { fprintf(stderr, "pffft_new_setup failed: %s\n", pffftb_last_error()); return 1; }   /* no GPU: no plan */
  float *x = (float *)pffft_aligned_malloc(batch * per * sizeof(float));    /* page-locked: DMA straight from here */
  float *y = (float *)pffft_aligned_malloc(batch * per * sizeof(float));
  float *z = (float *)pffft_aligned_malloc(batch * per * sizeof(float));
  if (!x || !y || !z) return 2;
  for (size_t i = 0; i < batch * per; ++i) x[i] = (float)rand() / (float)RAND_MAX * 2.f - 1.f;

  if (pffftb_transform_batch(s, x, y, batch, PFFFT_FORWARD, /*ordered=*/1) ||
      pffftb_transform_batch(s, y, z, batch, PFFFT_BACKWARD, 1)) {
    fprintf(stderr, "transform failed: %s\n", pffftb_last_error());
    return 3;
  }
  double worst = 0.0;
  for (size_t i = 0; i < batch * per; ++i) { const double e = fabs((double)z[i] / N - (double)x[i]); if (e > worst) worst = e; }
  printf("kernel %s on device %d: %zu x N=%d complex, max |ifft(fft(x))/N - x| = %.3g, %llu kernel launches\n",
         pffftb_setup_kernel(s), pffftb_setup_device(s), batch, N, worst, pffftb_launch_count());
  pffft_aligned_free(x); pffft_aligned_free(y); pffft_aligned_free(z);
  pffft_destroy_setup(s);
  return worst < 1e-5 ? 0 : 4;
}
