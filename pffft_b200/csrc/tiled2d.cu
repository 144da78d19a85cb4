// tiled2d.cu -- instantiations and launchers of the tiled two-dimensional large-N plan (tiled2d_kernels.cuh), float.
// Own translation unit so the C-ABI units stay small.  Default plan for 32768 and 65536 (PFFFT_B200_TILED2D=0|1 overrides).
#include <cuda_runtime.h>
#include <stdlib.h>
#include "internal_api.h"
#include "tiled2d_kernels.cuh"

namespace pf {
namespace {

template <int A1, int A2, int SIGN> struct T2DLaunch {
  using G = T2D<A1, A2>;
  static constexpr int MINB_A = 768 / (16 * A2), MINB_C = 768 / (16 * A1);       // 80 registers per thread (64 spills 16-24 of them)
  static constexpr size_t kSmemA = (size_t)16 * G::N2 * sizeof(cf), kSmemC = (size_t)16 * G::N1 * sizeof(cf);

  static int run(const cf* x, cf* S, cf* X, long long batch, const cf* tables, int sm_count, cudaStream_t st) {
    auto ka = k_t2d_A<float, A1, A2, SIGN, MINB_A>;
    auto kc = k_t2d_C<float, A1, A2, SIGN, MINB_C>;
    static thread_local int per_sm_a = 0, per_sm_c = 0;
    if (per_sm_a == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, ka, G::TA, kSmemA);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, kc, G::TC, kSmemC);
      if (per_sm_a < 1) per_sm_a = 1;
      if (per_sm_c < 1) per_sm_c = 1;
    }
    const cf* twA = tables;
    const cf* twC = twA + G::N2;
    const cf* tw2d = twC + G::N1;
    long long ga = batch * (G::N1 / 16), gc = batch * (G::N2 / 16);
    if (ga > (long long)sm_count * per_sm_a) ga = (long long)sm_count * per_sm_a;
    if (gc > (long long)sm_count * per_sm_c) gc = (long long)sm_count * per_sm_c;
    ka<<<(int)ga, G::TA, kSmemA, st>>>(x, S, batch, twA, tw2d);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    kc<<<(int)gc, G::TC, kSmemC, st>>>(S, X, batch, twC);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

}  // namespace

// (A1, A2) for a complex core, 0 when the size has no tiled plan
bool t2d_shape_for(int Nc, int* A1, int* A2) {
  switch (Nc) {
    case 16384: *A1 = 8; *A2 = 8; return true;
    case 32768: *A1 = 16; *A2 = 8; return true;
    case 65536: *A1 = 16; *A2 = 16; return true;
  }
  return false;
}
size_t t2d_table_cpx(int Nc) {
  int a1 = 0, a2 = 0;
  return t2d_shape_for(Nc, &a1, &a2) ? (size_t)16 * a1 + (size_t)16 * a2 + (size_t)Nc : 0;
}
void t2d_fill_tables_float(int Nc, float* dst) {
  switch (Nc) {
    case 16384: t2d_fill_tables<float, 8, 8>(dst); break;
    case 32768: t2d_fill_tables<float, 16, 8>(dst); break;
    case 65536: t2d_fill_tables<float, 16, 16>(dst); break;
  }
}
int t2d_launch_float(int Nc, int sign, const cf* x, cf* S, cf* X, long long batch, const cf* tables, int sm_count, cudaStream_t st) {
#define PF_T2D(nc, a1, a2) if (Nc == nc) return sign < 0 ? T2DLaunch<a1, a2, -1>::run(x, S, X, batch, tables, sm_count, st) \
                                                         : T2DLaunch<a1, a2, +1>::run(x, S, X, batch, tables, sm_count, st);
  PF_T2D(16384, 8, 8) PF_T2D(32768, 16, 8) PF_T2D(65536, 16, 16)
#undef PF_T2D
  set_error_msg("tiled2d: size not instantiated");
  return (int)cudaErrorInvalidValue;
}

}  // namespace pf
