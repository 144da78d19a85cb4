// radix_e.cu -- double-precision compile-time-radix CTA kernels, second translation unit: three-stage cores 576 ... 3840
#include "radix_impl.cuh"
namespace pf {
int radix_launch_double_e(int Nc, int lm, int sm, int sign, const double* in, double* out, long long batch, const cd* tw, const cd* twr,
                          int device, int sm_count, cudaStream_t st) {
  switch (Nc) {
    //                                          R1  R2  R3 TPC MINB
    case 576:  return radix_launch_modes<double, 9,  8,  8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 640:  return radix_launch_modes<double, 10, 8,  8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 720:  return radix_launch_modes<double, 10, 9,  8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 768:  return radix_launch_modes<double, 12, 8,  8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 800:  return radix_launch_modes<double, 10, 10, 8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 864:  return radix_launch_modes<double, 12, 9,  8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 960:  return radix_launch_modes<double, 12, 10, 8,  2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1152: return radix_launch_modes<double, 12, 12, 8,  1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1200: return radix_launch_modes<double, 12, 10, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1280: return radix_launch_modes<double, 16, 10, 8,  1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1440: return radix_launch_modes<double, 12, 12, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1600: return radix_launch_modes<double, 16, 10, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1728: return radix_launch_modes<double, 12, 12, 12, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1920: return radix_launch_modes<double, 16, 12, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    // 2160 ... 3840: radices up to 16 / 18 (a double radix-16 DFT holds 64 registers of data: one CTA per SM)
    case 2160: return radix_launch_modes<double, 12, 12, 15, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2304: return radix_launch_modes<double, 16, 12, 12, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2400: return radix_launch_modes<double, 16, 15, 10, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2560: return radix_launch_modes<double, 16, 16, 10, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2592: return radix_launch_modes<double, 9,  16, 18, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2880: return radix_launch_modes<double, 16, 15, 12, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3456: return radix_launch_modes<double, 16, 18, 12, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3600: return radix_launch_modes<double, 16, 15, 15, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 3840: return radix_launch_modes<double, 16, 16, 15, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return -1;
  }
}
}  // namespace pf
