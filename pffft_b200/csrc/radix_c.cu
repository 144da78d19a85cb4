// radix_c.cu -- compile-time-radix CTA kernels, third translation unit: the three-stage cores 1152 ... 3840 added in round 2b
#include "radix_impl.cuh"
namespace pf {
// Cores 1152 ... 14400 that had no tuned plan (two-launch split plans at 0.30-0.33 of the roofline or the generic shared-memory
// kernel): three stages of radices <= 16, no spills at three resident CTAs (round 2b, profiles/r02b_radix.md)
int radix_launch_float_d(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st) {
  switch (Nc) {
    //                                  R1  R2  R3  TPC MINB
    case 1152: return radix_launch_modes<float, 12, 12, 8,  2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1200: return radix_launch_modes<float, 12, 10, 10, 2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1280: return radix_launch_modes<float, 16, 10, 8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1440: return radix_launch_modes<float, 12, 12, 10, 2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1600: return radix_launch_modes<float, 16, 10, 10, 2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1728: return radix_launch_modes<float, 12, 12, 12, 2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1920: return radix_launch_modes<float, 16, 12, 10, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2304: return radix_launch_modes<float, 16, 12, 12, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3200: return radix_launch_modes<float, 20, 16, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3456: return radix_launch_modes<float, 16, 18, 12, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3600: return radix_launch_modes<float, 16, 15, 15, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3840: return radix_launch_modes<float, 16, 16, 15, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2160: return radix_launch_modes<float, 12, 12, 15, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2400: return radix_launch_modes<float, 16, 15, 10, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2880: return radix_launch_modes<float, 16, 15, 12, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    // 4320 ... 8000: radices up to 20 / 24 (two-launch split plans before)
    case 4320: return radix_launch_modes<float, 16, 18, 15, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 4608: return radix_launch_modes<float, 16, 16, 18, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 4800: return radix_launch_modes<float, 16, 20, 15, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 5184: return radix_launch_modes<float, 16, 18, 18, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 5760: return radix_launch_modes<float, 16, 18, 20, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 6400: return radix_launch_modes<float, 16, 20, 20, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 6912: return radix_launch_modes<float, 16, 18, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 7200: return radix_launch_modes<float, 15, 20, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 8000: return radix_launch_modes<float, 20, 20, 20, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    // 7680 / 9216: against the one-CTA split kernels (0.31 / 0.43) and the general-radix tiled plan (0.34 / 0.36); PFFFT_B200_RADIX_BIG=0 keeps those
    case 7680: return radix_launch_modes<float, 16, 20, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 9216: return radix_launch_modes<float, 16, 24, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    // 9600 ... 14400 (one CTA per SM, radices up to 30; two-launch split plans before)
    case 9600:  return radix_launch_modes<float, 20, 20, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 10800: return radix_launch_modes<float, 18, 20, 30, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 11520: return radix_launch_modes<float, 20, 24, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 12960: return radix_launch_modes<float, 18, 24, 30, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 13824: return radix_launch_modes<float, 24, 24, 24, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 14400: return radix_launch_modes<float, 24, 24, 25, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2560: return radix_launch_modes<float, 16, 16, 10, 1, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 5120: return radix_launch_modes<float, 16, 16, 20, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return -1;
  }
}
}  // namespace pf
