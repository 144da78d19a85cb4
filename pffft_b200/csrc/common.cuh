// common.cuh -- complex value type, small helpers shared by every kernel.
//
// Everything that is pure arithmetic/index math is `PF_HD` (host+device) so the very same
// code can be stepped lane-by-lane on the CPU by tests/emu (a development harness; the shipped
// library never executes these on the host).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PF_HD __host__ __device__ __forceinline__
#define PF_D  __device__ __forceinline__

namespace pf {

// (re, im) pair with vector alignment so global/shared accesses become LD/ST.64 (float) or .128 (double)
template <typename T> struct alignas(2 * sizeof(T)) cpx { T x, y; };
using cf = cpx<float>;
using cd = cpx<double>;

template <typename T> PF_HD cpx<T> mk(T x, T y) { cpx<T> r; r.x = x; r.y = y; return r; }

// ---- element-wise pair arithmetic.  On sm_100a a float pair lives in an aligned register pair and one FADD2 / FMUL2 /
// FFMA2 (add/mul/fma.rn.f32x2) works on both halves; its operands take a swap (.LO_HI), a half negation (.NP), a whole
// negation and a 32-bit broadcast (.F32 / immediate) for free, so x*(+-i), conj-like sign patterns and scalar twiddle
// parts cost nothing: a complex add is 1 instruction, a complex multiply 2, a twiddled butterfly 4 (scalar: 2 / 4 / 8).
// tools/ubench_fp32.cu measures the issue rates; PF_NO_PACKED_F32 switches back to scalar code (A/B builds).
#if defined(__CUDA_ARCH__) && !defined(PF_NO_PACKED_F32)
#define PF_PK 1
#else
#define PF_PK 0
#endif
#if PF_PK
// inline PTX rather than the __fadd2_rn/__fmul2_rn/__ffma2_rn intrinsics: same SASS (ptxas folds the mov.b64 packs into
// operand modifiers), but the intrinsics made cicc 8x slower on the big translation units (24 min for api_float.cu)
typedef unsigned long long pk64;
PF_D pk64 pk(cpx<float> a) { pk64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
PF_D cpx<float> upk(pk64 v) { cpx<float> r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
PF_D cpx<float> pk_add(cpx<float> a, cpx<float> b) { pk64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
PF_D cpx<float> pk_mul(cpx<float> a, cpx<float> b) { pk64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
PF_D cpx<float> pk_fma(cpx<float> a, cpx<float> b, cpx<float> c) {
  pk64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(r);
}
#endif
template <typename T> struct is_f32 { static constexpr bool value = false; };
template <> struct is_f32<float> { static constexpr bool value = true; };

template <typename T> PF_HD cpx<T> operator+(cpx<T> a, cpx<T> b) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_add(a, b);
#endif
  return mk<T>(a.x + b.x, a.y + b.y);
}
template <typename T> PF_HD cpx<T> operator-(cpx<T> a, cpx<T> b) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_add(a, mk<float>(-b.x, -b.y));
#endif
  return mk<T>(a.x - b.x, a.y - b.y);
}
template <typename T> PF_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
// a * s (both halves)
template <typename T> PF_HD cpx<T> scale(cpx<T> a, T s) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_mul(a, mk<float>(s, s));
#endif
  return mk<T>(a.x * s, a.y * s);
}
template <typename T> PF_HD cpx<T> scale2(cpx<T> a, T s) { return scale(a, s); }   // same, for scopes where `scale` is a variable
// a * s + c (both halves, fused)
template <typename T> PF_HD cpx<T> cfma(cpx<T> a, T s, cpx<T> c) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_fma(a, mk<float>(s, s), c);
#endif
  return mk<T>(fma(a.x, s, c.x), fma(a.y, s, c.y));
}
// (a.x*b.x + c.x, a.y*b.y + c.y) and (a.x*b.x, a.y*b.y)
template <typename T> PF_HD cpx<T> efma(cpx<T> a, cpx<T> b, cpx<T> c) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_fma(a, b, c);
#endif
  return mk<T>(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y));
}
template <typename T> PF_HD cpx<T> emul(cpx<T> a, cpx<T> b) {
#if PF_PK
  if constexpr (is_f32<T>::value) return pk_mul(a, b);
#endif
  return mk<T>(a.x * b.x, a.y * b.y);
}
// i*a and -i*a (register renaming in scalar code, an operand modifier in packed code)
template <typename T> PF_HD cpx<T> mul_pi(cpx<T> a) { return mk<T>(-a.y, a.x); }
template <typename T> PF_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }
// a * b = b.x*(a.x, a.y) + b.y*(-a.y, a.x)   (2 mul + 2 fma; packed: FMUL2 + FFMA2, b's halves as broadcast operands)
template <typename T> PF_HD cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return efma(mul_pi(a), mk<T>(b.y, b.y), emul(a, mk<T>(b.x, b.x)));
}
// a * (w.x + i*SIGNFLIP*w.y): SIGN=-1 keeps the table's forward sign, SIGN=+1 conjugates it:
//   a * conj(w) = w.x*(a.x, a.y) + w.y*(a.y, -a.x)
template <int SIGN, typename T> PF_HD cpx<T> cmul_dir(cpx<T> a, cpx<T> w) {
  if (SIGN < 0) return cmul(a, w);
  return efma(mul_mi(a), mk<T>(w.y, w.y), emul(a, mk<T>(w.x, w.x)));
}
// multiply by SIGN*i  (SIGN=-1 : *(-i) ; SIGN=+1 : *(+i))
template <int SIGN, typename T> PF_HD cpx<T> mul_si(cpx<T> a) {
  return SIGN < 0 ? mk<T>(a.y, -a.x) : mk<T>(-a.y, a.x);
}

// direction enum values are the reference's (include/pffft/pffft.h:108-117)
enum { DIR_FORWARD = 0, DIR_BACKWARD = 1 };
enum { XF_REAL = 0, XF_COMPLEX = 1 };

}  // namespace pf
