// radix_a.cu -- compile-time-radix CTA kernels, small cores (radix_kernels.cuh): 16 .. 432
#include "radix_impl.cuh"
namespace pf {
int radix_launch_float_a(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st) {
  switch (Nc) {
    //                                  R1  R2  R3  TPC MINB
    case 16:  return radix_launch_modes<4,  4,  1,  64, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 48:  return radix_launch_modes<16, 3,  1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 80:  return radix_launch_modes<16, 5,  1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 144: return radix_launch_modes<12, 12, 1,  20, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 240: return radix_launch_modes<16, 15, 1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 400: return radix_launch_modes<20, 20, 1,  12, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 432: return radix_launch_modes<12, 12, 3,  2,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return -1;
  }
}
}  // namespace pf
