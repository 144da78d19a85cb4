// radix_b.cu -- compile-time-radix CTA kernels, larger cores (radix_kernels.cuh): 1296 .. 12000
#include <stdlib.h>
#include "radix_impl.cuh"
namespace pf {
int radix_launch_float_a(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st);
int radix_launch_float_d(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st);
int radix_launch_float_x(int alt, int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st);
bool radix_core_supported(int Nc, const char** name) {
  static const struct { int nc; const char* name; } k[] = {
      {16, "radix_4x4"}, {48, "radix_8x6"}, {80, "radix_10x8"}, {144, "radix_12x12"}, {240, "radix_16x15"}, {400, "radix_20x20"},
      {432, "radix_24x18"}, {720, "radix_30x24"}, {1296, "radix_12x12x9"}, {2000, "radix_25x10x8"}, {2592, "radix_9x16x18"}, {4000, "radix_25x16x10"},
      {6000, "radix_15x20x20"}, {12000, "radix_25x24x20"},
      {1152, "radix_12x12x8"}, {1200, "radix_12x10x10"}, {1280, "radix_16x10x8"}, {1440, "radix_12x12x10"}, {1600, "radix_16x10x10"},
      {1728, "radix_12x12x12"}, {1920, "radix_16x12x10"}, {2304, "radix_16x12x12"}, {3200, "radix_20x16x10"}, {3456, "radix_16x18x12"},
      {3600, "radix_16x15x15"}, {3840, "radix_16x16x15"}, {2160, "radix_12x12x15"}, {2400, "radix_16x15x10"}, {2880, "radix_16x15x12"},
      {4320, "radix_16x18x15"}, {4608, "radix_16x16x18"}, {4800, "radix_16x20x15"}, {5184, "radix_16x18x18"}, {5760, "radix_16x18x20"},
      {6400, "radix_16x20x20"}, {6912, "radix_16x18x24"}, {7200, "radix_15x20x24"}, {8000, "radix_20x20x20"}, {7680, "radix_16x20x24"}, {9216, "radix_16x24x24"}, {2560, "radix_16x16x10"}, {5120, "radix_16x16x20"},
      {9600, "radix_20x20x24"}, {10800, "radix_18x20x30"}, {11520, "radix_20x24x24"}, {12960, "radix_18x24x30"}, {13824, "radix_24x24x24"}, {14400, "radix_24x24x25"}};
  static const bool big = !(getenv("PFFFT_B200_RADIX_BIG") && atoi(getenv("PFFFT_B200_RADIX_BIG")) == 0);
  if (!big && (Nc == 7680 || Nc == 9216 || Nc == 2560 || Nc == 5120)) return false;
  for (const auto& e : k) if (e.nc == Nc) { if (name) *name = e.name; return true; }
  return false;
}
int radix_launch_float(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                       int device, int sm_count, cudaStream_t st) {
  static const int alt = getenv("PFFFT_B200_RADIX_ALT") ? atoi(getenv("PFFFT_B200_RADIX_ALT")) : 0;      // A/B shapes (radix_x.cu)
  if (alt > 0) { const int rc = radix_launch_float_x(alt, Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st); if (rc != -2 && rc != -1) return rc; }
  // Stage shapes of the cores 2000 ... 6000: measured against three alternatives each (radix_x.cu, profiles/r02b_radix.md):
  // radices and their ORDER decide how many threads a transform gets (Nc / smallest radix) and with it the register budget.
  switch (Nc) {
    case 1152: case 1200: case 1280: case 1440: case 1600: case 1728: case 1920: case 2304: case 3200: case 3456: case 3600: case 3840:
    case 2160: case 2400: case 2880: case 4320: case 4608: case 4800: case 5184: case 5760: case 6400: case 6912: case 7200: case 8000: case 7680: case 9216: case 2560: case 5120:
    case 9600: case 10800: case 11520: case 12960: case 13824: case 14400:
      return radix_launch_float_d(Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 1296:  return radix_launch_modes<float, 12, 12, 9,  2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2000:  return radix_launch_modes<float, 25, 10, 8,  1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2592:  return radix_launch_modes<float, 9,  16, 18, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 4000:  return radix_launch_modes<float, 25, 16, 10, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 6000:  return radix_launch_modes<float, 15, 20, 20, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 12000: return radix_launch_modes<float, 25, 24, 20, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return radix_launch_float_a(Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  }
}
}  // namespace pf
