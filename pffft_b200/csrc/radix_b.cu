// radix_b.cu -- compile-time-radix CTA kernels, larger cores (radix_kernels.cuh): 1296 .. 12000
#include <stdlib.h>
#include "radix_impl.cuh"
namespace pf {
int radix_launch_float_a(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st);
int radix_launch_float_c(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                         int device, int sm_count, cudaStream_t st);
bool radix_core_supported(int Nc, const char** name) {
  static const struct { int nc; const char* name; } k[] = {
      {16, "radix_4x4"}, {48, "radix_16x3"}, {80, "radix_16x5"}, {144, "radix_12x12"}, {240, "radix_16x15"}, {400, "radix_20x20"},
      {432, "radix_12x12x3"}, {1296, "radix_12x12x9"}, {2000, "radix_20x10x10"}, {2592, "radix_18x12x12"}, {4000, "radix_20x20x10"},
      {6000, "radix_20x20x15"}, {12000, "radix_25x24x20"}};
  for (const auto& e : k) if (e.nc == Nc) { if (name) *name = e.name; return true; }
  return false;
}
int radix_launch_float(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cf* tw, const cf* twr,
                       int device, int sm_count, cudaStream_t st) {
  static const bool one_cta = getenv("PFFFT_B200_RADIX_MINB1") && atoi(getenv("PFFFT_B200_RADIX_MINB1")) == 1;
  if (one_cta && (Nc == 2000 || Nc == 2592 || Nc == 4000 || Nc == 6000))
    return radix_launch_float_c(Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  switch (Nc) {
    case 1296:  return radix_launch_modes<12, 12, 9,  2, 3>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2000:  return radix_launch_modes<20, 10, 10, 2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 2592:  return radix_launch_modes<18, 12, 12, 2, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 4000:  return radix_launch_modes<20, 20, 10, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 6000:  return radix_launch_modes<20, 20, 15, 1, 2>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    case 12000: return radix_launch_modes<25, 24, 20, 1, 1>(lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
    default: return radix_launch_float_a(Nc, lm, sm, sign, in, out, batch, tw, twr, device, sm_count, st);
  }
}
}  // namespace pf
