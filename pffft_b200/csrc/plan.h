// plan.h -- host-side plan algebra (no CUDA): size validation, factorisation, twiddle tables.
//
// Restates the *contract* of pffft_new_setup (ref src/pffft_priv_impl.h:1062-1112) and the size
// helpers (:78-114, src/pffft_common.c:47-55); the factor order and table contents are this
// engine's own (Stockham stages over one exp(-2 pi i k/Nc) table, computed in long double and
// rounded once -- the reference fills FFTPACK `wa` arrays with float cosf/sinf, :948-993).
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>

namespace pfplan {

enum { kSimd = 4 };                 // layout granularity kept from the reference (pffft_simd_size)
enum { kMaxN = 1 << 26 };           // ref pffft_priv_impl.h:1069

inline int min_fft_size(int transform) {      // ref :78-89   (0 = REAL, 1 = COMPLEX)
  if (transform == 0) return 2 * kSimd * kSimd;
  if (transform == 1) return kSimd * kSimd;
  return 1;
}
inline int is_valid_size(int N, int transform) {   // ref :91-98
  const int nmin = min_fft_size(transform);
  int r = N;
  while (r >= 5 * nmin && (r % 5) == 0) r /= 5;
  while (r >= 3 * nmin && (r % 3) == 0) r /= 3;
  while (r >= 2 * nmin && (r % 2) == 0) r /= 2;
  return r == nmin ? 1 : 0;
}
inline int nearest_transform_size(int N, int transform, int higher) {   // ref :100-114
  const int nmin = min_fft_size(transform);
  if (N < nmin) N = nmin;
  const int d = higher ? nmin : -nmin;
  N = higher ? nmin * ((N + nmin - 1) / nmin) : nmin * (N / nmin);
  for (;; N += d)
    if (is_valid_size(N, transform)) return N;
}
inline int next_power_of_two(int N) {     // ref pffft_common.c:24-37: smallest 2^k >= N in 32-bit unsigned arithmetic
  unsigned v = (unsigned)N - 1u;
  for (unsigned sh = 1; sh < 32; sh <<= 1) v |= v >> sh;
  return (int)(v + 1u);
}
inline int is_power_of_two(int N) { return (N != 0 && (N & (N - 1)) == 0) ? 1 : 0; }  // ref pffft_common.c:39-43

// new_setup's acceptance rule (ref :1066-1078, :1105-1109)
inline bool setup_size_ok(int N, int transform) {
  if (N <= 0 || N > kMaxN) return false;
  if (transform != 0 && transform != 1) return false;
  if (N % min_fft_size(transform)) return false;
  int r = N / kSimd;
  for (int p : {2, 3, 5}) while (r % p == 0) r /= p;
  return r == 1;
}

// radix list for the Stockham stages of an Nc-point complex core: 4s first, one 2 if needed, then 3s, 5s
inline std::vector<int> factorize(int Nc) {
  std::vector<int> f;
  int n2 = 0, n3 = 0, n5 = 0;
  while (Nc % 2 == 0) { Nc /= 2; ++n2; }
  while (Nc % 3 == 0) { Nc /= 3; ++n3; }
  while (Nc % 5 == 0) { Nc /= 5; ++n5; }
  if (Nc != 1) return {};
  for (int i = 0; i < n2 / 2; ++i) f.push_back(4);
  if (n2 & 1) f.push_back(2);
  for (int i = 0; i < n3; ++i) f.push_back(3);
  for (int i = 0; i < n5; ++i) f.push_back(5);
  return f;
}

// (cos, sin)(-2 pi k / n) with octant reduction in long double, so tables are correctly rounded
inline void unit_root(long long k, long long n, long double* c, long double* s) {
  k %= n;
  const long long q = (4 * k) / n, r = 4 * k - q * n;   // quadrant, remainder: angle = (q + r/n) * pi/2
  const long double hp = 1.57079632679489661923132169163975144L;
  long double c0, s0;
  if (r == 0) { c0 = 1; s0 = 0; }
  else if (2 * r == n) { c0 = s0 = 0.70710678118654752440084436210484904L; }
  else if (2 * r < n) { long double a = hp * (long double)r / (long double)n; c0 = cosl(a); s0 = sinl(a); }
  else { long double a = hp * (long double)(n - r) / (long double)n; c0 = sinl(a); s0 = cosl(a); }
  long double cc, ss;
  switch (q) { case 0: cc = c0; ss = s0; break; case 1: cc = -s0; ss = c0; break;
               case 2: cc = -c0; ss = -s0; break; default: cc = s0; ss = -c0; break; }
  *c = cc; *s = -ss;   // forward sign
}
template <typename T> inline void fill_roots(T* dst /*2*count*/, long long count, long long n) {
  for (long long k = 0; k < count; ++k) {
    long double c, s;
    unit_root(k, n, &c, &s);
    dst[2 * k] = (T)c; dst[2 * k + 1] = (T)s;
  }
}

// ---- overlap-save block algebra of pffastconv_apply
struct BlockPlan {            // the reference's loop (pffastconv.c:156-167 / :204-210) in closed form
  long long n_full = 0;       // blocks with procLen == Nfft
  int stride = 0;             // outputs (= input advance) per full block
  long long tail_off = -1;    // offset of the one partial block (flush only), -1 if none
  int tail_out = 0;           // its output count
  long long produced = 0;     // total outputs = final inpOff
};

inline BlockPlan plan_blocks(long long inputLen, int Nfft, int filterLen, int flush, bool even_out) {
  BlockPlan bp;
  int s = Nfft - filterLen + 1;
  if (even_out) s &= ~1;
  bp.stride = s;
  if (s <= 0) return bp;
  const long long maxOff = flush ? (inputLen - filterLen + 1) : (inputLen - Nfft + 1);
  if (maxOff <= 0) return bp;
  if (inputLen >= Nfft) bp.n_full = (inputLen - Nfft) / s + 1;
  long long off = bp.n_full * s;
  if (off < maxOff) {                      // only reachable with flush: remaining input shorter than Nfft
    const long long procLen = inputLen - off;
    long long nout = procLen - filterLen + 1;
    if (even_out) nout &= ~1LL;
    if (nout > 0) { bp.tail_off = off; bp.tail_out = (int)nout; off += nout; }
  }
  bp.produced = off;
  return bp;
}

}  // namespace pfplan
