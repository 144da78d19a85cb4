// butterfly.cuh -- the arithmetic core: small-radix DFTs and fully-unrolled register FFTs.
//
// Replaces (does not translate) the reference's passf2/3/4/5_ps (src/pffft_priv_impl.h:122-321)
// and, through the N/2 complex packing, radf*/radb* (:323-807).  The reference walks FFTPACK
// passes over 4-lane SIMD vectors; here a thread owns R points in registers, runs a complete
// radix-R DIT network on them with compile-time twiddles (FMA-folded butterflies), and exchanges
// data with other lanes only between such networks.
//
// Sign convention (include/pffft/pffft.h:134): SIGN=-1 is PFFFT_FORWARD
// (e^{-2 pi i jk/N}), SIGN=+1 is PFFFT_BACKWARD; nothing is normalised.
#pragma once
#include <math.h>
#include "common.cuh"

namespace pf {

// ---------------------------------------------------------------------------------------------
// compile-time trigonometry: cos/sin(2*pi*j/n) evaluated exactly-rounded enough for double
// (quadrant/half-quadrant reduction on the integer fraction, then Taylor on |x| <= pi/4).
// ---------------------------------------------------------------------------------------------
namespace ct {
constexpr double kPi = 3.14159265358979323846264338327950288;
__host__ __device__ constexpr double taylor_sin(double x) {
  double x2 = x * x, term = x, sum = x;
  for (int k = 1; k <= 12; ++k) { term *= -x2 / double((2 * k) * (2 * k + 1)); sum += term; }
  return sum;
}
__host__ __device__ constexpr double taylor_cos(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int k = 1; k <= 12; ++k) { term *= -x2 / double((2 * k - 1) * (2 * k)); sum += term; }
  return sum;
}
struct cs { double c, s; };
// (cos, sin) of 2*pi*j/n for 0 <= j < n
__host__ __device__ constexpr cs cossin2pi(long long j, long long n) {
  j %= n; if (j < 0) j += n;
  long long q = (4 * j) / n;            // quadrant
  long long r = 4 * j - q * n;          // position inside the quadrant, angle = (r/n) * pi/2
  double c0 = 1.0, s0 = 0.0;
  if (2 * r <= n) { double a = (double(r) / double(n)) * (kPi / 2); c0 = taylor_cos(a); s0 = taylor_sin(a); }
  else { double a = (double(n - r) / double(n)) * (kPi / 2); c0 = taylor_sin(a); s0 = taylor_cos(a); }
  if (r == 0) { c0 = 1.0; s0 = 0.0; }
  if (2 * r == n) { c0 = 0.70710678118654752440; s0 = c0; }
  switch (q) {
    case 0: return cs{c0, s0};
    case 1: return cs{-s0, c0};
    case 2: return cs{-c0, -s0};
    default: return cs{s0, -c0};
  }
}
__host__ __device__ constexpr int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
__host__ __device__ constexpr int bitrev(int x, int bits) { int r = 0; for (int i = 0; i < bits; ++i) if (x & (1 << i)) r |= 1 << (bits - 1 - i); return r; }
}  // namespace ct

// ---------------------------------------------------------------------------------------------
// radix-R DFT on R values held in an array (generic kernels; runtime-selected radix).
// c_k = sum_j a_j * exp(SIGN * 2 pi i * j k / R), in place, natural order in and out.
// Constants as in SURVEY App. B (taur/taui, tr11..ti12) -- they are just cos/sin of 2pi/3, 2pi/5.
// ---------------------------------------------------------------------------------------------
template <int SIGN, typename T> PF_HD void dft2(cpx<T>* a) {
  cpx<T> t = a[0] - a[1]; a[0] = a[0] + a[1]; a[1] = t;
}
template <int SIGN, typename T> PF_HD void dft3(cpx<T>* a) {
  const T hs3 = T(SIGN) * T(0.86602540378443864676372317075294);
  const cpx<T> t1 = a[1] + a[2];
  const cpx<T> m = cfma(t1, T(-0.5), a[0]);      // a0 - (a1+a2)/2
  const cpx<T> id = mul_pi(a[1] - a[2]);         // i*(a1-a2)
  a[0] = a[0] + t1;
  a[1] = cfma(id, hs3, m);                       // m + i*SIGN*sin(2pi/3)*(a1-a2)
  a[2] = cfma(id, -hs3, m);
}
template <int SIGN, typename T> PF_HD void dft4(cpx<T>* a) {
  cpx<T> t0 = a[0] + a[2], t1 = a[0] - a[2], t2 = a[1] + a[3];
  cpx<T> t3 = mul_si<SIGN>(a[1] - a[3]);
  a[0] = t0 + t2; a[2] = t0 - t2; a[1] = t1 + t3; a[3] = t1 - t3;
}
template <int SIGN, typename T> PF_HD void dft5(cpx<T>* a) {
  const T tr11 = T(0.30901699437494742410229341718282), ti11 = T(SIGN) * T(0.95105651629515357211643933337938);
  const T tr12 = T(-0.80901699437494742410229341718282), ti12 = T(SIGN) * T(0.58778525229247312916870595463907);
  const cpx<T> t1 = a[1] + a[4], t2 = a[2] + a[3];
  const cpx<T> i3 = mul_pi(a[1] - a[4]), i4 = mul_pi(a[2] - a[3]);      // i*(a1-a4), i*(a2-a3)
  const cpx<T> m1 = cfma(t2, tr12, cfma(t1, tr11, a[0]));
  const cpx<T> m2 = cfma(t2, tr11, cfma(t1, tr12, a[0]));
  const cpx<T> n1 = cfma(i4, ti12, scale(i3, ti11));                     // i*(ti11 t3 + ti12 t4)
  const cpx<T> n2 = cfma(i4, -ti11, scale(i3, ti12));                    // i*(ti12 t3 - ti11 t4)
  a[0] = a[0] + t1 + t2;
  a[1] = m1 + n1; a[4] = m1 - n1;
  a[2] = m2 + n2; a[3] = m2 - n2;
}
template <int R, int SIGN, typename T> PF_HD void dftR(cpx<T>* a) {
  if (R == 2) dft2<SIGN>(a);
  else if (R == 3) dft3<SIGN>(a);
  else if (R == 4) dft4<SIGN>(a);
  else dft5<SIGN>(a);
}

// ---------------------------------------------------------------------------------------------
// Register FFT: N = 2^m points in cpx<T> v[N], decimation in time, all indices compile-time.
// Input convention: v[p] holds x[bitrev(p)]  (the caller chooses registers at load time, so the
// bit reversal costs nothing); output v[k] = X[k] in natural order.
// One butterfly (A,B) <- (A + w B, A - w B), w = exp(SIGN 2 pi i J/G), G = group size:
//   J=0, G/4      : adds only
//   G/8, 3G/8     : 2 adds + 4 fma          (w = (+-1 + SIGN i)/sqrt2)
//   general       : 8 fma                    (each output its own 2-fma chain)
// -> 428 instructions for N=32 (456 for the mul-then-add form), every product fused.
// ---------------------------------------------------------------------------------------------
template <int J, int G, int SIGN, typename T> PF_HD void dit_bfly(cpx<T>& A, cpx<T>& B) {
  const cpx<T> a = A, b = B;
  if constexpr (J == 0) {
    A = a + b; B = a - b;
  } else if constexpr (4 * J == G) {
    const cpx<T> t = mul_si<SIGN>(b);
    A = a + t; B = a - t;
  } else if constexpr (8 * J == G) {             // w = (1 + SIGN i)/sqrt2:  w b = h (b + SIGN i b)
    const T h = T(0.70710678118654752440084436210485);
    const cpx<T> s = b + mul_si<SIGN>(b);
    A = cfma(s, h, a); B = cfma(s, -h, a);
  } else if constexpr (8 * J == 3 * G) {         // w = (-1 + SIGN i)/sqrt2: w b = -h (b - SIGN i b)
    const T h = T(0.70710678118654752440084436210485);
    const cpx<T> s = b - mul_si<SIGN>(b);
    A = cfma(s, -h, a); B = cfma(s, h, a);
  } else {
    constexpr ct::cs w = ct::cossin2pi(J, G);
    const T c = T(w.c), s = T(SIGN) * T(w.s);
    // w b = c b + s (i b); both outputs as their own 2-fma chains (8 fma per butterfly, 4 FFMA2 packed).  The 6-fma
    // variant B = 2a - A re-injects A's rounding error into B, which costs ~4 dB of spur-free range on pure tones
    // (tests/test_pffft.c wants >= 140 dB).
    const cpx<T> ib = mul_pi(b);
    A = cfma(ib, s, cfma(b, c, a));
    B = cfma(ib, -s, cfma(b, -c, a));
  }
}

// one DIT level: groups of size G over the N-point array (element p -> v[BASE + p*STRIDE])
template <int N, int G, int SIGN, int BASE, int STRIDE, int P = 0, typename T, int TOTAL>
PF_HD void dit_level(cpx<T> (&v)[TOTAL]) {
  if constexpr (P < N / 2) {
    constexpr int H = G / 2;
    constexpr int g = P / H, j = P % H;
    dit_bfly<j, G, SIGN>(v[BASE + (g * G + j) * STRIDE], v[BASE + (g * G + j + H) * STRIDE]);
    dit_level<N, G, SIGN, BASE, STRIDE, P + 1>(v);
  }
}
template <int N, int SIGN, int BASE, int STRIDE, int G = 2, typename T, int TOTAL>
PF_HD void dit_fft(cpx<T> (&v)[TOTAL]) {
  if constexpr (G <= N) {
    dit_level<N, G, SIGN, BASE, STRIDE>(v);
    dit_fft<N, SIGN, BASE, STRIDE, 2 * G>(v);
  }
}
// convenience: whole array
template <int N, int SIGN, typename T> PF_HD void reg_fft(cpx<T> (&v)[N]) { dit_fft<N, SIGN, 0, 1>(v); }


// ---------------------------------------------------------------------------------------------
// Small mixed-radix register DFTs (R = 3,5,6,9,10,12,15,18,20,24,25,27,30) for the non-power-of-two warp kernels:
// (nested) Cooley-Tukey splits R = R1 x R2 with compile-time twiddles, natural order in and out.
//   X[k1 + R1*k2] = sum_n2 W_R2^{n2 k2} * ( W_R^{n2 k1} * sum_n1 a[R2*n1 + n2] W_R1^{n1 k1} )
// These are the register-resident counterparts of the reference's passf3_ps / passf5_ps
// (src/pffft_priv_impl.h:151-183, :256-321).
// ---------------------------------------------------------------------------------------------
template <int J, int N, int SIGN, typename T> PF_HD cpx<T> mul_root(cpx<T> x) {   // x * exp(SIGN 2 pi i J/N)
  if constexpr (J % N == 0) return x;
  else {
    constexpr ct::cs w = ct::cossin2pi(J, N);
    const T c = T(w.c), s = T(SIGN) * T(w.s);
    return cfma(mul_pi(x), s, scale(x, c));      // c x + s (i x)
  }
}
template <int R, int SIGN, typename T> PF_HD void dft_small(cpx<T>* a);   // any supported size, defined below
template <int R1, int R2, int SIGN, int I = 0, typename T> PF_HD void ct_twiddle(cpx<T>* b) {   // b[k1*R2 + n2] *= W_R^{n2 k1}
  if constexpr (I < R1 * R2) {
    constexpr int k1 = I / R2, n2 = I % R2;
    b[I] = mul_root<k1 * n2, R1 * R2, SIGN>(b[I]);
    ct_twiddle<R1, R2, SIGN, I + 1>(b);
  }
}
template <int R1, int R2, int SIGN, typename T> PF_HD void dft_ct(cpx<T>* a) {
  cpx<T> b[R1 * R2];
#pragma unroll
  for (int n2 = 0; n2 < R2; ++n2) {              // R1-point DFTs down the columns
    cpx<T> t[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) t[n1] = a[R2 * n1 + n2];
    dft_small<R1, SIGN>(t);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) b[k1 * R2 + n2] = t[k1];
  }
  ct_twiddle<R1, R2, SIGN>(b);
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {              // R2-point DFTs along the rows
    cpx<T> t[R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) t[n2] = b[k1 * R2 + n2];
    dft_small<R2, SIGN>(t);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) a[k1 + R1 * k2] = t[k2];
  }
}
template <int R, int SIGN, typename T> PF_HD void dft_small(cpx<T>* a) {
  if constexpr (R == 2) dft2<SIGN>(a);
  else if constexpr (R == 3) dft3<SIGN>(a);
  else if constexpr (R == 4) dft4<SIGN>(a);
  else if constexpr (R == 5) dft5<SIGN>(a);
  else if constexpr (R == 1) { }
  else if constexpr (R == 6) dft_ct<2, 3, SIGN>(a);
  else if constexpr (R == 8) dft_ct<2, 4, SIGN>(a);
  else if constexpr (R == 9) dft_ct<3, 3, SIGN>(a);
  else if constexpr (R == 10) dft_ct<2, 5, SIGN>(a);
  else if constexpr (R == 12) dft_ct<4, 3, SIGN>(a);
  else if constexpr (R == 15) dft_ct<3, 5, SIGN>(a);
  else if constexpr (R == 16) dft_ct<4, 4, SIGN>(a);
  else if constexpr (R == 18) dft_ct<2, 9, SIGN>(a);
  else if constexpr (R == 20) dft_ct<4, 5, SIGN>(a);
  else if constexpr (R == 24) dft_ct<4, 6, SIGN>(a);
  else if constexpr (R == 25) dft_ct<5, 5, SIGN>(a);
  else if constexpr (R == 27) dft_ct<3, 9, SIGN>(a);
  else { static_assert(R == 30, "dft_small: unsupported size"); dft_ct<5, 6, SIGN>(a); }
}

}  // namespace pf
