/* pffft_oracle_impl.h -- body of the CPU restatement, included once per precision by pffft_oracle.c.
 *
 * TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the product
 * (pffft_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * What it restates (all file:line into /root/reference, marton78/pffft @ a4b0359):
 *   - plan acceptance + size helpers          src/pffft_priv_impl.h:76-114, :1062-1112, decompose :904-928
 *   - complex transform = 4 polyphase N/4-point FFTs (the 4 SIMD lanes, :1490-1495; cfftf1_ps :1004-1048)
 *     + cross-lane radix-4 "finalize"/"preprocess" with the e[] twiddles (:1089-1097, :1195-1270)
 *   - real transform: the same N = 4 x N/4 decomposition applied to the real signal; the reference
 *     specialises the sub-transforms to half-complex FFTPACK passes (rfftf1/rfftb1 + radf/radb, :323-901,
 *     real_finalize/preprocess :1273-1462) -- here they are evaluated with the complex routine on the
 *     Hermitian-extended data, which is the same arithmetic function; canonical packing per pffft.h:144-155
 *   - z-domain layout / pffft_zreorder        :1158-1193 (+ reversed_copy :1125-1139)
 *   - zconvolve_accumulate / _no_accu         :1534-1684, unfused mul/add order of src/simd/pf_float.h:76
 *   - pffastconv_new_setup/apply              src/pffastconv.c:58-263
 * Scalar code, no SIMD: per-element results are NOT bit-identical to the reference's (different summation
 * order); parity is pinned by tests/test_oracle.py against oracle/_ref (the unmodified reference compiled
 * here) and against the committed golden vectors: relmax <= 2e-6 float / 1e-13 double (pow2).
 *
 * Macros expected: T (scalar type), PFX(name) (symbol prefix), SETUP_T, TCOS/TSIN.
 */

typedef struct { T re, im; } PFX(cpx);

struct SETUP_T {
  int N, Ncvec, transform;
  int nsub;               /* N/4: length of each polyphase sub-transform */
  int nfac, fac[32];      /* factors of nsub, reference preference order (:915-921, :932, :965) */
  PFX(cpx) *roots;        /* exp(-2 pi i k/nsub), k < nsub            (twiddle[] analogue, :981-993) */
  PFX(cpx) *e;            /* exp(-2 pi i j k/N), j=1..3, k < nsub     (e[] analogue, :1089-1097) */
};

static int PFX(decompose)(int n, int *fac, const int *order) {      /* ref :904-928 */
  int nf = 0, j, i;
  for (j = 0; order[j]; ++j) {
    const int r = order[j];
    while (n != 1 && n % r == 0) {
      fac[nf++] = r; n /= r;
      if (r == 2 && nf != 1) { for (i = nf - 1; i > 0; --i) fac[i] = fac[i - 1]; fac[0] = 2; }   /* a lone 2 goes first */
    }
  }
  return n == 1 ? nf : -1;
}

int PFX(simd_size)(void) { return 4; }
const char *PFX(simd_arch)(void) { return "oracle-scalar"; }
int PFX(min_fft_size)(int transform) { return transform == 0 ? 32 : (transform == 1 ? 16 : 1); }   /* ref :78-89 */
int PFX(is_valid_size)(int N, int transform) {                                                       /* ref :91-98 */
  const int nmin = PFX(min_fft_size)(transform);
  int r = N;
  while (r >= 5 * nmin && r % 5 == 0) r /= 5;
  while (r >= 3 * nmin && r % 3 == 0) r /= 3;
  while (r >= 2 * nmin && r % 2 == 0) r /= 2;
  return r == nmin;
}
int PFX(nearest_transform_size)(int N, int transform, int higher) {                                  /* ref :100-114 */
  const int nmin = PFX(min_fft_size)(transform);
  int d = higher ? nmin : -nmin;
  if (N < nmin) N = nmin;
  N = higher ? nmin * ((N + nmin - 1) / nmin) : nmin * (N / nmin);
  for (;; N += d) if (PFX(is_valid_size)(N, transform)) return N;
}
int PFX(next_power_of_two)(int N) { return oracle_next_pow2(N); }
int PFX(is_power_of_two)(int N) { return N && !(N & (N - 1)); }
void *PFX(aligned_malloc)(size_t nb) { return oracle_aligned_malloc(nb); }
void PFX(aligned_free)(void *p) { oracle_aligned_free(p); }

void PFX(destroy_setup)(struct SETUP_T *s) { if (!s) return; free(s->roots); free(s->e); free(s); }

struct SETUP_T *PFX(new_setup)(int N, int transform) {               /* ref :1062-1112 */
  static const int order_c[] = {5, 3, 4, 2, 0}, order_r[] = {4, 2, 3, 5, 0};
  struct SETUP_T *s;
  int k, j;
  if (N < 0 || N > (1 << 26)) return NULL;
  if (transform == 0) { if (N % 32 || N <= 0) return NULL; }
  else if (transform == 1) { if (N % 16 || N <= 0) return NULL; }
  else return NULL;
  s = (struct SETUP_T *)calloc(1, sizeof(*s));
  s->N = N; s->transform = transform; s->Ncvec = (transform == 0 ? N / 2 : N) / 4; s->nsub = N / 4;
  s->nfac = PFX(decompose)(s->nsub, s->fac, transform == 0 ? order_r : order_c);
  if (s->nfac < 0) { free(s); return NULL; }                         /* other prime factors -> NULL (:1105-1109) */
  s->roots = (PFX(cpx) *)malloc(sizeof(PFX(cpx)) * (size_t)s->nsub);
  s->e = (PFX(cpx) *)malloc(sizeof(PFX(cpx)) * 3 * (size_t)s->nsub);
  for (k = 0; k < s->nsub; ++k) {
    const T a = (T)k * ((T)(2 * 3.14159265358979323846264338327950288) / (T)s->nsub);   /* trig in working precision, like :956-957 */
    s->roots[k].re = TCOS(a); s->roots[k].im = -TSIN(a);
    for (j = 1; j < 4; ++j) {
      const T A = -2 * (T)3.14159265358979323846264338327950288 * (T)j * (T)k / (T)N;   /* :1093 */
      s->e[(j - 1) * s->nsub + k].re = TCOS(A); s->e[(j - 1) * s->nsub + k].im = TSIN(A);
    }
  }
  return s;
}

/* out[0..n) = DFT_n of in[0], in[stride], ...; sign=-1 forward.  Decimation in time over the factor list. */
static void PFX(fft_rec)(const struct SETUP_T *s, int n, int stride, const PFX(cpx) *in, PFX(cpx) *out,
                         const int *fac, int sign) {
  int r, m, j, k, q;
  if (n == 1) { out[0] = in[0]; return; }
  r = fac[0]; m = n / r;
  for (j = 0; j < r; ++j) PFX(fft_rec)(s, m, stride * r, in + (size_t)j * stride, out + (size_t)j * m, fac + 1, sign);
  for (k = 0; k < m; ++k) {
    PFX(cpx) t[5];
    for (j = 0; j < r; ++j) {                       /* twiddle W_n^{jk} (passf*: VCPLXMUL by wa, :137,:174,:234) */
      const PFX(cpx) w = s->roots[(size_t)((long long)j * k % n) * (s->nsub / n)];
      const T wr = w.re, wi = sign < 0 ? w.im : -w.im;
      const PFX(cpx) v = out[(size_t)j * m + k];
      t[j].re = v.re * wr - v.im * wi; t[j].im = v.re * wi + v.im * wr;
    }
    for (q = 0; q < r; ++q) {                       /* radix-r butterfly: sum_j t_j W_r^{jq} */
      T ar = 0, ai = 0;
      for (j = 0; j < r; ++j) {
        const PFX(cpx) w = s->roots[(size_t)(((long long)j * q) % r) * (s->nsub / r)];
        const T wr = w.re, wi = sign < 0 ? w.im : -w.im;
        ar += t[j].re * wr - t[j].im * wi; ai += t[j].re * wi + t[j].im * wr;
      }
      out[(size_t)q * m + k].re = ar; out[(size_t)q * m + k].im = ai;
    }
  }
}

/* complex N-point transform, canonical in / canonical out (unnormalised) */
static void PFX(cplx_core)(const struct SETUP_T *s, const PFX(cpx) *x, PFX(cpx) *X, int sign) {
  const int ns = s->nsub, N = s->N;
  PFX(cpx) *lane = (PFX(cpx) *)malloc(sizeof(PFX(cpx)) * (size_t)N * 2);
  PFX(cpx) *sub = lane + N;
  int j, k, q;
  if (sign < 0) {
    for (j = 0; j < 4; ++j) PFX(fft_rec)(s, ns, 4, x + j, sub + (size_t)j * ns, s->fac, -1);   /* polyphase lanes */
    for (k = 0; k < ns; ++k) {                                   /* finalize (:1195-1237) */
      PFX(cpx) t[4];
      t[0] = sub[k];
      for (j = 1; j < 4; ++j) {
        const PFX(cpx) w = s->e[(size_t)(j - 1) * ns + k], v = sub[(size_t)j * ns + k];
        t[j].re = v.re * w.re - v.im * w.im; t[j].im = v.re * w.im + v.im * w.re;
      }
      for (q = 0; q < 4; ++q) {                                   /* * (-i)^{jq} */
        T ar = t[0].re, ai = t[0].im;
        for (j = 1; j < 4; ++j) {
          switch ((j * q) & 3) {
            case 0: ar += t[j].re; ai += t[j].im; break;
            case 1: ar += t[j].im; ai -= t[j].re; break;
            case 2: ar -= t[j].re; ai -= t[j].im; break;
            default: ar -= t[j].im; ai += t[j].re; break;
          }
        }
        X[(size_t)q * ns + k].re = ar; X[(size_t)q * ns + k].im = ai;
      }
    }
  } else {
    for (k = 0; k < ns; ++k) {                                   /* preprocess (:1239-1270) */
      for (j = 0; j < 4; ++j) {
        T ar = 0, ai = 0;
        for (q = 0; q < 4; ++q) {                                 /* * (+i)^{jq} */
          const PFX(cpx) v = x[(size_t)q * ns + k];
          switch ((j * q) & 3) {
            case 0: ar += v.re; ai += v.im; break;
            case 1: ar -= v.im; ai += v.re; break;
            case 2: ar -= v.re; ai -= v.im; break;
            default: ar += v.im; ai -= v.re; break;
          }
        }
        if (j) {
          const PFX(cpx) w = s->e[(size_t)(j - 1) * ns + k];      /* conj(e) */
          const T br = ar * w.re + ai * w.im, bi = ai * w.re - ar * w.im;
          ar = br; ai = bi;
        }
        lane[(size_t)j * ns + k].re = ar; lane[(size_t)j * ns + k].im = ai;
      }
    }
    for (j = 0; j < 4; ++j) {
      PFX(fft_rec)(s, ns, 1, lane + (size_t)j * ns, sub + (size_t)j * ns, s->fac, +1);
      for (k = 0; k < ns; ++k) X[4 * (size_t)k + j] = sub[(size_t)j * ns + k];
    }
  }
  free(lane);
}

/* z-domain position (scalar index of the real part; imaginary part 4 further) of internal element
 * (quarter q, index u): memory is blocks of 8 four-lane vectors r0 i0 r1 i1 r2 i2 r3 i3 (:1158-1193) */
static size_t PFX(zidx)(int q, int u) { return 32 * (size_t)(u >> 2) + 8 * (size_t)q + (size_t)(u & 3); }
/* canonical slot held by internal element (q,u) */
static int PFX(zbin)(const struct SETUP_T *s, int q, int u) {
  if (s->transform == 1) return q * (s->N / 4) + u;                           /* kk = k/4 + (k%4)*Ncvec/4, :1181-1185 */
  {
    const int n8 = s->N / 8;
    switch (q) {
      case 0: return u;                                                       /* INTERLEAVE2 of vin[8k+0,1], :1166 */
      case 2: return 2 * n8 + u;                                              /* INTERLEAVE2 of vin[8k+4,5], :1167 */
      case 1: return u == 0 ? n8 : 2 * n8 - u;                                /* reversed_copy(vin+2) below N/4, :1169 */
      default: return u == 0 ? 3 * n8 : 4 * n8 - u;                           /* reversed_copy(vin+6) below N/2, :1170 */
    }
  }
}
void PFX(zreorder)(struct SETUP_T *s, const T *in, T *out, int direction) {
  const int nu = (s->transform == 1 ? s->N : s->N / 2) / 4;
  int q, u;
  for (q = 0; q < 4; ++q)
    for (u = 0; u < nu; ++u) {
      const size_t z = PFX(zidx)(q, u);
      const size_t c = 2 * (size_t)PFX(zbin)(s, q, u);
      if (direction == 0) { out[c] = in[z]; out[c + 1] = in[z + 4]; }
      else { out[z] = in[c]; out[z + 4] = in[c + 1]; }
    }
}

static void PFX(run)(struct SETUP_T *s, const T *in, T *out, int direction, int ordered) {
  const int N = s->N;
  const size_t per = s->transform == 0 ? (size_t)N : 2 * (size_t)N;
  PFX(cpx) *a = (PFX(cpx) *)malloc(sizeof(PFX(cpx)) * (size_t)N * 2), *b = a + N;
  T *tmp = (T *)malloc(sizeof(T) * per);
  int k;
  if (s->transform == 1) {
    if (direction == 0) {
      memcpy(a, in, sizeof(T) * per);
      PFX(cplx_core)(s, a, b, -1);
      if (ordered) memcpy(out, b, sizeof(T) * per); else { memcpy(tmp, b, sizeof(T) * per); PFX(zreorder)(s, tmp, out, 1); }
    } else {
      if (ordered) memcpy(a, in, sizeof(T) * per); else { PFX(zreorder)(s, in, tmp, 0); memcpy(a, tmp, sizeof(T) * per); }
      PFX(cplx_core)(s, a, b, +1);
      memcpy(out, b, sizeof(T) * per);
    }
  } else {
    if (direction == 0) {
      for (k = 0; k < N; ++k) { a[k].re = in[k]; a[k].im = 0; }
      PFX(cplx_core)(s, a, b, -1);
      tmp[0] = b[0].re; tmp[1] = b[N / 2].re;                               /* slot 0 = (DC, Nyquist), pffft.h:144-155 */
      for (k = 1; k < N / 2; ++k) { tmp[2 * k] = b[k].re; tmp[2 * k + 1] = b[k].im; }
      if (ordered) memcpy(out, tmp, sizeof(T) * per); else { T *t2 = (T *)malloc(sizeof(T) * per); memcpy(t2, tmp, sizeof(T) * per); PFX(zreorder)(s, t2, out, 1); free(t2); }
    } else {
      if (ordered) memcpy(tmp, in, sizeof(T) * per); else PFX(zreorder)(s, in, tmp, 0);
      a[0].re = tmp[0]; a[0].im = 0; a[N / 2].re = tmp[1]; a[N / 2].im = 0;
      for (k = 1; k < N / 2; ++k) { a[k].re = tmp[2 * k]; a[k].im = tmp[2 * k + 1]; a[N - k].re = tmp[2 * k]; a[N - k].im = -tmp[2 * k + 1]; }
      PFX(cplx_core)(s, a, b, +1);
      for (k = 0; k < N; ++k) out[k] = b[k].re;
    }
  }
  free(tmp); free(a);
}
void PFX(transform)(struct SETUP_T *s, const T *in, T *out, T *work, int direction) { (void)work; PFX(run)(s, in, out, direction, 0); }
void PFX(transform_ordered)(struct SETUP_T *s, const T *in, T *out, T *work, int direction) { (void)work; PFX(run)(s, in, out, direction, 1); }

/* ref :1534-1684: per 4-lane pair ab (+)= (a*b)*scaling with separately rounded products and sums */
static void PFX(zconv)(struct SETUP_T *s, const T *a, const T *b, T *ab, T scaling, int acc) {
  const size_t per = s->transform == 0 ? (size_t)s->N : 2 * (size_t)s->N;
  const T ar0 = a[0], ai0 = a[4], br0 = b[0], bi0 = b[4], abr0 = ab[0], abi0 = ab[4];
  size_t g; int l;
  for (g = 0; g < per; g += 8)
    for (l = 0; l < 4; ++l) {
      const T ar = a[g + l], ai = a[g + 4 + l], br = b[g + l], bi = b[g + 4 + l];
      volatile T p0 = ar * br, p1 = ai * bi, p2 = ai * br, p3 = ar * bi;   /* volatile: forbid fma contraction */
      volatile T re = p0 - p1, im = p2 + p3;
      volatile T sr = re * scaling, si = im * scaling;
      if (acc) { ab[g + l] = sr + ab[g + l]; ab[g + 4 + l] = si + ab[g + 4 + l]; }
      else { ab[g + l] = sr; ab[g + 4 + l] = si; }
    }
  if (s->transform == 0) {                                               /* DC / Nyquist are independent reals, :1626-1629 */
    volatile T d0 = ar0 * br0, d1 = ai0 * bi0;
    volatile T e0 = d0 * scaling, e1 = d1 * scaling;
    ab[0] = acc ? abr0 + e0 : e0; ab[4] = acc ? abi0 + e1 : e1;
  }
}
void PFX(zconvolve_accumulate)(struct SETUP_T *s, const T *a, const T *b, T *ab, T sc) { PFX(zconv)(s, a, b, ab, sc, 1); }
void PFX(zconvolve_no_accu)(struct SETUP_T *s, const T *a, const T *b, T *ab, T sc) { PFX(zconv)(s, a, b, ab, sc, 0); }
