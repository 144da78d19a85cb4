// radix_a.cu -- compile-time-radix CTA kernels, small cores (radix_kernels.cuh): 16 .. 432
#include <stdlib.h>
#include "radix_impl.cuh"
namespace pf {
int radix_launch_float_a(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st) {
  // Resident CTAs per SM, measured (profiles/r02b_radix.md; PFFFT_B200_RADIX_VAR=0 forces the round-2 shapes, =2 the
  // one-more-CTA shapes everywhere): the small cores 48 / 80 / 144 gain 2-5 % from a fourth CTA (64 registers), the backward
  // real transforms of the core 400 27 % from a third (0.33 -> 0.42); 240 and the forward transforms of 400 lose 8-18 %.
  static const int var = getenv("PFFFT_B200_RADIX_VAR") ? atoi(getenv("PFFFT_B200_RADIX_VAR")) : -1;
  const bool bwd_real = lm == L_R_ORD || lm == L_R_Z;
  const bool more = var == 2 || (var < 0 && (Nc == 48 || Nc == 80 || Nc == 144 || (Nc == 400 && bwd_real)));
  // cores 48 and 80 with BALANCED radices (8 x 6, 10 x 8: both stages keep 6-10 of the transform's 8-10 threads busy; 16 x 3 and
  // 16 x 5 left 3 resp. 5 of 16 threads working in the first stage).  PFFFT_B200_RADIX_BAL=0: the round-2 shapes.
  static const bool balanced = !(getenv("PFFFT_B200_RADIX_BAL") && atoi(getenv("PFFFT_B200_RADIX_BAL")) == 0);
  if (balanced && Nc == 48) return radix_launch_modes<float, 8,  6, 1, 32, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  if (balanced && Nc == 80) return radix_launch_modes<float, 10, 8, 1, 24, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  if (more) switch (Nc) {
    case 48:  return radix_launch_modes<float, 16, 3,  1,  16, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 80:  return radix_launch_modes<float, 16, 5,  1,  16, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 144: return radix_launch_modes<float, 12, 12, 1,  20, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 240: if (var == 2) return radix_launch_modes<float, 16, 15, 1,  16, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st); break;
    case 400: return radix_launch_modes<float, 20, 20, 1,  12, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 432: if (var == 2) return radix_launch_modes<float, 12, 12, 3,  2,  4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st); break;
    default: break;
  }
  switch (Nc) {
    //                                  R1  R2  R3  TPC MINB
    case 16:  return radix_launch_modes<float, 4,  4,  1,  64, 4>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 48:  return radix_launch_modes<float, 16, 3,  1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 80:  return radix_launch_modes<float, 16, 5,  1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 144: return radix_launch_modes<float, 12, 12, 1,  20, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 240: return radix_launch_modes<float, 16, 15, 1,  16, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 400: return radix_launch_modes<float, 20, 20, 1,  12, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 432:                                                    // two stages (24 x 18) since round 2b; PFFFT_B200_RADIX_432=3: the three-stage 12 x 12 x 3
      if (getenv("PFFFT_B200_RADIX_432") && atoi(getenv("PFFFT_B200_RADIX_432")) == 3)
        return radix_launch_modes<float, 12, 12, 3,  2,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
      return radix_launch_modes<float, 24, 18, 1,  10, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 720: return radix_launch_modes<float, 30, 24, 1,  8,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);   // round 2b (was on the generic kernel: 0.36)
    default: return -1;
  }
}
}  // namespace pf
