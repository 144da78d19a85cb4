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
template <typename T> PF_HD cpx<T> operator+(cpx<T> a, cpx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> PF_HD cpx<T> operator-(cpx<T> a, cpx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> PF_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
template <typename T> PF_HD cpx<T> scale(cpx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <typename T> PF_HD cpx<T> scale2(cpx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }   // same, for scopes where `scale` is a variable
// a * b (contraction to FMA allowed: 2 mul + 2 fma)
template <typename T> PF_HD cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return mk<T>(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// a * (b.x + i*SIGNFLIP*b.y): SIGN=-1 keeps the table's forward sign, SIGN=+1 conjugates it
template <int SIGN, typename T> PF_HD cpx<T> cmul_dir(cpx<T> a, cpx<T> w) {
  if (SIGN < 0) return cmul(a, w);
  return mk<T>(fma(a.x, w.x, a.y * w.y), fma(a.y, w.x, -(a.x * w.y)));
}
// multiply by SIGN*i  (SIGN=-1 : *(-i) ; SIGN=+1 : *(+i))
template <int SIGN, typename T> PF_HD cpx<T> mul_si(cpx<T> a) {
  return SIGN < 0 ? mk<T>(a.y, -a.x) : mk<T>(-a.y, a.x);
}

// direction enum values are the reference's (include/pffft/pffft.h:108-117)
enum { DIR_FORWARD = 0, DIR_BACKWARD = 1 };
enum { XF_REAL = 0, XF_COMPLEX = 1 };

}  // namespace pf
