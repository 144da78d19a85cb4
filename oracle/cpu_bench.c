/* cpu_bench.c -- TEST/BENCH INFRASTRUCTURE (never linked into the product).
 *
 * Times a pffft-ABI CPU library (the unmodified reference built into oracle/_ref, or this repo's
 * restatement) on the host cores: `nthreads` POSIX threads share ONE read-only PFFFT_Setup (legal:
 * include/pffft/pffft.h:102-106 of the reference) and each transforms its contiguous slice of a
 * DRAM-resident batch with its own work buffer, calling the library one vector at a time exactly
 * as a reference user would.  Wall-clock via clock_gettime(CLOCK_MONOTONIC).
 *
 * The library is opened with dlopen so the same driver serves both checkers.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void* (*new_setup_fn)(int, int);
typedef void (*destroy_fn)(void*);
typedef void (*xform_fn)(void*, const float*, float*, float*, int);
typedef void* (*amalloc_fn)(size_t);
typedef void (*afree_fn)(void*);

typedef struct {
  xform_fn xform; void* setup;
  float* in; float* out; float* work;
  size_t per; long first, count; int fwd_inv; int iters; unsigned seed;
} job_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* first touch by the thread that will use the slice (NUMA placement), uniform(-1,1) from a per-thread xorshift */
static void* filler(void* p) {
  job_t* j = (job_t*)p;
  uint32_t st = j->seed * 2654435761u + (uint32_t)j->first * 40503u + 1u;
  for (size_t i = (size_t)j->first * j->per; i < (size_t)(j->first + j->count) * j->per; ++i) {
    st ^= st << 13; st ^= st >> 17; st ^= st << 5;
    j->in[i] = (float)((st >> 8) * (1.0 / 8388608.0) - 1.0);
    j->out[i] = 0.f;
  }
  return NULL;
}

static void* worker(void* p) {
  job_t* j = (job_t*)p;
  for (int it = 0; it < j->iters; ++it)
    for (long b = j->first; b < j->first + j->count; ++b) {
      j->xform(j->setup, j->in + b * j->per, j->out + b * j->per, j->work, 0);
      if (j->fwd_inv) j->xform(j->setup, j->out + b * j->per, j->out + b * j->per, j->work, 1);
    }
  return NULL;
}

/* returns seconds for `iters` passes over `batch` transforms (forward, or forward+backward when
 * fwd_inv != 0) with `nthreads` threads; < 0 on error.  ordered != 0 -> pffft_transform_ordered. */
double cpu_bench_transform(const char* libpath, int N, int transform, long batch, int nthreads, int iters,
                           int fwd_inv, int ordered, unsigned seed) {
  void* h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "cpu_bench: dlopen(%s): %s\n", libpath, dlerror()); return -1.0; }
  new_setup_fn new_setup = (new_setup_fn)dlsym(h, "pffft_new_setup");
  destroy_fn destroy = (destroy_fn)dlsym(h, "pffft_destroy_setup");
  xform_fn xform = (xform_fn)dlsym(h, ordered ? "pffft_transform_ordered" : "pffft_transform");
  amalloc_fn amalloc = (amalloc_fn)dlsym(h, "pffft_aligned_malloc");
  afree_fn afree = (afree_fn)dlsym(h, "pffft_aligned_free");
  if (!new_setup || !destroy || !xform || !amalloc || !afree) return -2.0;
  void* s = new_setup(N, transform);
  if (!s) return -3.0;
  const size_t per = transform == 0 ? (size_t)N : 2 * (size_t)N;
  float* in = (float*)amalloc(per * batch * sizeof(float));
  float* out = (float*)amalloc(per * batch * sizeof(float));
  if (!in || !out) return -4.0;
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    long lo = batch * t / nthreads, hi = batch * (t + 1) / nthreads;
    jobs[t] = (job_t){xform, s, in, out, (float*)amalloc(per * sizeof(float)), per, lo, hi - lo, fwd_inv, iters, seed ? seed : 1u};
  }
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, filler, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  /* warm-up pass (page faults, caches), then the timed passes */
  { job_t* warm = (job_t*)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; ++t) { warm[t] = jobs[t]; warm[t].iters = 1; pthread_create(&th[t], NULL, worker, &warm[t]); }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(warm); }
  const double t0 = now_s();
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  const double dt = now_s() - t0;
  for (int t = 0; t < nthreads; ++t) afree(jobs[t].work);
  free(jobs); free(th); afree(in); afree(out); destroy(s);
  dlclose(h);
  return dt;
}

/* pffastconv on one stream, single thread (a PFFASTCONV_Setup is not shareable, pffastconv.h:77-81):
 * returns seconds per pffastconv_apply(len samples, flush=1); *produced receives the output count. */
typedef void* (*fc_new_fn)(const float*, int, int*, int);
typedef int (*fc_apply_fn)(void*, const float*, int, float*, int);
typedef void (*fc_destroy_fn)(void*);
double cpu_bench_fastconv(const char* libpath, int len, int taps, int iters, int* produced) {
  void* h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1.0;
  fc_new_fn fnew = (fc_new_fn)dlsym(h, "pffastconv_new_setup");
  fc_apply_fn fapply = (fc_apply_fn)dlsym(h, "pffastconv_apply");
  fc_destroy_fn fdel = (fc_destroy_fn)dlsym(h, "pffastconv_destroy_setup");
  if (!fnew || !fapply || !fdel) return -2.0;
  float* x = (float*)malloc(sizeof(float) * (size_t)len);
  float* y = (float*)malloc(sizeof(float) * (size_t)len);
  float* hc = (float*)malloc(sizeof(float) * (size_t)taps);
  static const float pat[3] = {-1.f, 1.f, 0.5f};
  for (int i = 0; i < len; ++i) x[i] = (float)(i % 4093);
  for (int j = 0; j < taps; ++j) hc[j] = pat[j % 3];
  int bl = 0;
  void* s = fnew(hc, taps, &bl, 0);
  if (!s) return -3.0;
  int n = fapply(s, x, len, y, 1);
  const double t0 = now_s();
  for (int it = 0; it < iters; ++it) n = fapply(s, x, len, y, 1);
  const double dt = (now_s() - t0) / (iters > 0 ? iters : 1);
  if (produced) *produced = n;
  fdel(s); free(x); free(y); free(hc); dlclose(h);
  return dt;
}
