/* pffft_oracle.c -- plain-C restatement of the reference's hot path (see pffft_oracle_impl.h for the
 * per-function citations).  TEST INFRASTRUCTURE: built into oracle/liboracle.so by oracle/Makefile and
 * used only as a checker by tests/, __graft_entry__.smoke() and bench.py's CPU legs.  The product
 * (pffft_b200/) never links, loads or calls it; there is no CPU path in the product.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this restatement against the unmodified reference
 * compiled into oracle/_ref/ (every valid size class, both precisions, ordered and z-domain, zreorder and
 * zconvolve bit-exact, pffastconv lengths + values) and against tests/golden/pffft_golden.npz.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int oracle_next_pow2(int N) {            /* ref src/pffft_common.c:24-37 */
  unsigned v = (unsigned)N - 1u, sh;
  for (sh = 1; sh < 32; sh <<= 1) v |= v >> sh;
  return (int)(v + 1u);
}
static void *oracle_aligned_malloc(size_t nb) { /* ref src/pffft_common.c:12-22: 64-byte aligned, raw pointer stashed in front */
  void *raw = malloc(nb + 64 + sizeof(void *));
  uintptr_t u;
  if (!raw) return NULL;
  u = ((uintptr_t)raw + 63 + sizeof(void *)) & ~(uintptr_t)63;
  ((void **)u)[-1] = raw;
  return (void *)u;
}
static void oracle_aligned_free(void *p) { if (p) free(((void **)p)[-1]); }

#define PASTE_(a, b) a##b
#define PASTE(a, b) PASTE_(a, b)

/* ---- float instantiation: pffft_* ---- */
#define T float
#define PFX(n) PASTE(pffft_, n)
#define SETUP_T PFFFT_Setup
#define TCOS cosf
#define TSIN sinf
#include "pffft_oracle_impl.h"
#undef T
#undef PFX
#undef SETUP_T
#undef TCOS
#undef TSIN

/* ---- double instantiation: pffftd_* ---- */
#define T double
#define PFX(n) PASTE(pffftd_, n)
#define SETUP_T PFFFTD_Setup
#define TCOS cos
#define TSIN sin
#include "pffft_oracle_impl.h"
#undef T
#undef PFX
#undef SETUP_T
#undef TCOS
#undef TSIN

/* ---- overlap-save convolution, ref src/pffastconv.c:58-263 (float only, like the reference) ---- */
struct PFFASTCONV_Setup {
  float *Xt, *Xf, *Hf, *Mf;
  struct PFFFT_Setup *st;
  int filterLen, Nfft, flags;
  float scale;
};
enum { FC_CPLX_INP_OUT = 1, FC_CPLX_FILTER = 2, FC_SINGLE_FFT = 16, FC_CORRELATION = 64 };

void *pffastconv_malloc(size_t nb) { return oracle_aligned_malloc(nb); }
void pffastconv_free(void *p) { oracle_aligned_free(p); }
int pffastconv_simd_size(void) { return 4; }

struct PFFASTCONV_Setup *pffastconv_new_setup(const float *h, int filterLen, int *blockLen, int flags) {   /* ref :58-116 */
  const int cf = ((flags & FC_CPLX_INP_OUT) && (flags & FC_SINGLE_FFT)) ? 2 : 1;
  int Nfft = 2 * oracle_next_pow2(filterLen - 1), i;
  struct PFFASTCONV_Setup *s;
  if (Nfft < 32) Nfft = 32;
  if (flags & FC_CPLX_FILTER) return NULL;
  if (*blockLen > Nfft) Nfft = oracle_next_pow2(*blockLen);
  *blockLen = Nfft;
  Nfft *= cf;
  s = (struct PFFASTCONV_Setup *)calloc(1, sizeof(*s));
  s->Xt = (float *)calloc((size_t)Nfft, sizeof(float)); s->Xf = (float *)calloc((size_t)Nfft, sizeof(float));
  s->Hf = (float *)calloc((size_t)Nfft, sizeof(float)); s->Mf = (float *)calloc((size_t)Nfft, sizeof(float));
  s->st = pffft_new_setup(Nfft, 0);
  s->filterLen = cf == 2 ? 2 * filterLen - 1 : filterLen;
  s->Nfft = Nfft; s->flags = flags; s->scale = (float)(1.0 / Nfft);
  for (i = 0; i < filterLen; ++i)
    s->Xt[(Nfft - cf * i) & (Nfft - 1)] = (flags & FC_CORRELATION) ? h[i] : h[filterLen - 1 - i];
  pffft_transform(s->st, s->Xt, s->Hf, s->Mf, 0);
  return s;
}
void pffastconv_destroy_setup(struct PFFASTCONV_Setup *s) {
  if (!s) return;
  pffft_destroy_setup(s->st); free(s->Xt); free(s->Xf); free(s->Hf); free(s->Mf); free(s);
}
static void fc_block(struct PFFASTCONV_Setup *s) {           /* one block: FFT, x Hf/Nfft, inverse (ref :235-254) */
  pffft_transform(s->st, s->Xt, s->Xf, s->Mf, 0);
  pffft_zconvolve_no_accu(s->st, s->Xf, s->Hf, s->Mf, s->scale);
  pffft_transform(s->st, s->Mf, s->Xf, s->Xt, 1);
}
int pffastconv_apply(struct PFFASTCONV_Setup *s, const float *X, int cplxInputLen, float *Y, int applyFlush) {   /* ref :133-263 */
  const int Nfft = s->Nfft, F = s->filterLen, flags = s->flags;
  const int cf = ((flags & FC_CPLX_INP_OUT) && (flags & FC_SINGLE_FFT)) ? 2 : 1;
  const int inputLen = cf * cplxInputLen;
  const int maxOff = applyFlush ? (inputLen - F + 1) : (inputLen - Nfft + 1);
  const int parts = (cf == 1 && (flags & FC_CPLX_INP_OUT)) ? 2 : 1;
  int off, numOut = 0, procLen, j, part;
  for (off = 0; off < maxOff; off += numOut) {
    procLen = (inputLen - off) >= Nfft ? Nfft : (inputLen - off);
    numOut = procLen - F + 1;
    if (cf == 2) { numOut &= ~1; if (!numOut) break; }
    for (part = 0; part < parts; ++part) {
      const int step = parts, base = parts * off + part;     /* de-interleave re / im streams (ref :212-224) */
      for (j = 0; j < procLen; ++j) s->Xt[j] = X[base + step * j];
      for (; j < Nfft; ++j) s->Xt[j] = 0.f;
      fc_block(s);
      for (j = 0; j < numOut; ++j) Y[base + step * j] = s->Xf[j];
    }
  }
  return off / cf;
}
