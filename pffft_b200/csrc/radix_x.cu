// radix_x.cu -- alternative stage shapes of the larger radix cores, forward complex / forward real only, for A/B runs
// (PFFFT_B200_RADIX_ALT=1..3; profiles/r02b_radix.md).  Shapes that win move into radix_b.cu / radix_c.cu with all modes.
#include <stdlib.h>
#include "radix_impl.cuh"
namespace pf {
template <int R1, int R2, int R3, int TPC, int MINB>
static int alt2(int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr, int device, int sm_count, cudaStream_t st) {
  if (lm == L_C_ORD && sm == S_C_ORD && sign < 0) return radix_launch_one<float, R1, R2, R3, L_C_ORD, S_C_ORD, -1, TPC, MINB>(in, out, batch, tw, twr, device, sm_count, st);
  if (lm == L_R_TIME && sm == S_R_ORD && sign < 0) return radix_launch_one<float, R1, R2, R3, L_R_TIME, S_R_ORD, -1, TPC, MINB>(in, out, batch, tw, twr, device, sm_count, st);
  return -1;
}
int radix_launch_float_x(int alt, int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st) {
#define A(n, nc, ...) if (alt == n && Nc == nc) return alt2<__VA_ARGS__>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  //            R1  R2  R3 TPC MINB
  A(1, 2000,  10, 20, 10, 2, 1)   A(2, 2000,  16, 25, 5,  1, 1)   A(3, 2000,  25, 10, 8,  1, 2)
  A(1, 2592,  16, 18, 9,  1, 2)   A(2, 2592,  12, 12, 18, 1, 2)   A(3, 2592,  9,  16, 18, 1, 2)
  A(1, 4000,  16, 25, 10, 1, 1)   A(2, 4000,  10, 20, 20, 1, 1)   A(3, 4000,  25, 16, 10, 1, 1)
  A(1, 6000,  16, 25, 15, 1, 1)   A(2, 6000,  15, 20, 20, 1, 1)   A(3, 6000,  25, 16, 15, 1, 1)
  A(1, 12000, 16, 25, 30, 1, 1)   A(2, 12000, 30, 25, 16, 1, 1)   A(3, 12000, 20, 24, 25, 1, 1)
#undef A
  return -2;                                                       // no such alternative: the caller uses the default shape
}
}  // namespace pf
