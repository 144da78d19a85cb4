// radix_c.cu -- compile-time-radix CTA kernels, the larger cores built for ONE resident CTA per SM (more registers, no spills):
// A/B against the two-CTA builds of radix_b.cu (PFFFT_B200_RADIX_MINB1=1 selects these), see profiles/r02b_radix.md
#include "radix_impl.cuh"
namespace pf {
int radix_launch_float_c(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st) {
  switch (Nc) {
    case 2000:  return radix_launch_modes<20, 10, 10, 2, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2592:  return radix_launch_modes<18, 12, 12, 2, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 4000:  return radix_launch_modes<20, 20, 10, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 6000:  return radix_launch_modes<20, 20, 15, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return -1;
  }
}
}  // namespace pf
