// radix_d.cu -- compile-time-radix CTA kernels in DOUBLE precision: the cores below 512 and the mixed-radix cores up to 4000
// that the double 16x16xC kernels (512 ... 4096) do not cover and that ran on the generic shared-memory kernel (0.37-0.47 of
// the roofline).  Radices <= 12 per thread (a double radix-12 DFT holds 48 registers of data); shapes balanced like the
// float ones (profiles/r02b_radix.md).  PFFFT_B200_RADIX_D=0 switches them off.
#include <stdlib.h>
#include "radix_impl.cuh"
namespace pf {
namespace {
const struct { int nc; const char* name; } kCoresD[] = {
    {16, "radix_4x4"}, {32, "radix_8x4"}, {48, "radix_8x6"}, {64, "radix_8x8"}, {80, "radix_10x8"}, {96, "radix_12x8"},
    {128, "radix_8x4x4"}, {144, "radix_12x12"}, {160, "radix_8x5x4"}, {192, "radix_8x6x4"}, {240, "radix_8x6x5"},
    {256, "radix_8x8x4"}, {288, "radix_8x6x6"}, {320, "radix_8x8x5"}, {384, "radix_8x8x6"}, {400, "radix_10x10x4"},
    {432, "radix_9x8x6"}, {480, "radix_10x8x6"}, {1296, "radix_12x12x9"}, {2000, "radix_10x20x10"},
    {576, "radix_9x8x8"}, {640, "radix_10x8x8"}, {720, "radix_10x9x8"}, {768, "radix_12x8x8"}, {800, "radix_10x10x8"}, {864, "radix_12x9x8"},
    {960, "radix_12x10x8"}, {1152, "radix_12x12x8"}, {1200, "radix_12x10x10"}, {1280, "radix_16x10x8"}, {1440, "radix_12x12x10"},
    {1600, "radix_16x10x10"}, {1728, "radix_12x12x12"}, {1920, "radix_16x12x10"},
    {2160, "radix_12x12x15"}, {2304, "radix_16x12x12"}, {2400, "radix_16x15x10"}, {2560, "radix_16x16x10"}, {2592, "radix_9x16x18"},
    {2880, "radix_16x15x12"}, {3456, "radix_16x18x12"}, {3600, "radix_16x15x15"}, {3840, "radix_16x16x15"}};
}
int radix_launch_double_e(int Nc, int lm, int sm, int sign, const double* in, double* out, long long batch, const cd* tw, const cd* twr,
                          int device, int sm_count, cudaStream_t st);
bool radix_core_supported_double(int Nc, const char** name) {
  static const bool on = !(getenv("PFFFT_B200_RADIX_D") && atoi(getenv("PFFFT_B200_RADIX_D")) == 0);
  if (!on) return false;
  for (const auto& e : kCoresD) if (e.nc == Nc) { if (name) *name = e.name; return true; }
  return false;
}
int radix_launch_double(int Nc, int lm, int sm, int sign, const double* in, double* out, long long batch, const cd* tw, const cd* twr,
                        int device, int sm_count, cudaStream_t st) {
  switch (Nc) {
    //                                          R1  R2  R3 TPC MINB
    case 16:   return radix_launch_modes<double, 4,  4,  1, 64, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 32:   return radix_launch_modes<double, 8,  4,  1, 32, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 48:   return radix_launch_modes<double, 8,  6,  1, 32, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 64:   return radix_launch_modes<double, 8,  8,  1, 32, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 80:   return radix_launch_modes<double, 10, 8,  1, 24, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 96:   return radix_launch_modes<double, 12, 8,  1, 20, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 128:  return radix_launch_modes<double, 8,  4,  4, 8,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 144:  return radix_launch_modes<double, 12, 12, 1, 20, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 160:  return radix_launch_modes<double, 8,  5,  4, 6,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 192:  return radix_launch_modes<double, 8,  6,  4, 5,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 240:  return radix_launch_modes<double, 8,  6,  5, 4,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 256:  return radix_launch_modes<double, 8,  8,  4, 4,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 288:  return radix_launch_modes<double, 8,  6,  6, 4,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 320:  return radix_launch_modes<double, 8,  8,  5, 4,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 384:  return radix_launch_modes<double, 8,  8,  6, 4,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 400:  return radix_launch_modes<double, 10, 10, 4, 2,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 432:  return radix_launch_modes<double, 9,  8,  6, 3,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 480:  return radix_launch_modes<double, 10, 8,  6, 3,  2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1296: return radix_launch_modes<double, 12, 12, 9, 1,  3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2000: return radix_launch_modes<double, 10, 20, 10, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return radix_launch_double_e(Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  }
}
}  // namespace pf
