// tiled2d_general.cu -- instantiations and launchers of the general-radix tiled two-dimensional plan (tiled2d_kernels.cuh):
// float complex cores Nc = 256*A1*A2.  NOT YET RUN ON HARDWARE (written after the GPU budget of round 1 was spent; verified by
// CPU stepping, tests/test_host_logic.py): reached only with PFFFT_B200_TILED2D_GENERAL=1.  Own translation unit.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <vector>
#include "internal_api.h"
#include "tiled2d_kernels.cuh"

namespace pf {
namespace {

template <typename T, int A1, int A2, int SIGN> struct T2DGLaunch {
  using G = T2D<A1, A2>;
  using cx = cpx<T>;
  static constexpr int MINB = sizeof(T) == 4 ? 3 : 1;                             // float: 80 registers, 768 threads per SM
  static constexpr size_t kSmemA = (size_t)16 * G::N2 * sizeof(cx), kSmemC = (size_t)16 * G::N1 * sizeof(cx);
  static int run(const cx* x, cx* S, cx* X, long long batch, const cx* tables, int sm_count, cudaStream_t st) {
    auto ka = k_t2dg_A<T, A1, A2, SIGN, MINB>;
    auto kc = k_t2dg_C<T, A1, A2, SIGN, MINB>;
    static PerDeviceInt occ_a, occ_c;
    const int dev = current_device();
    int arc = 0;
    const int per_sm_a = occ_a.get(dev, [&]() -> int {
      if (kSmemA > 48 * 1024) arc = (int)cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemA);
      if (arc) return -1;
      int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka, 256, kSmemA); return n < 1 ? 1 : n; });
    const int per_sm_c = occ_c.get(dev, [&]() -> int {
      if (kSmemC > 48 * 1024) arc = (int)cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemC);
      if (arc) return -1;
      int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kc, 256, kSmemC); return n < 1 ? 1 : n; });
    if (arc) { set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", (cudaError_t)arc); return arc; }
    const cx* twA = tables;
    const cx* twC = twA + G::N2;
    const cx* tw2d = twC + G::N1;
    long long ga = batch * (G::N1 / 16), gc = batch * (G::N2 / 16);
    if (ga > (long long)sm_count * per_sm_a) ga = (long long)sm_count * per_sm_a;
    if (gc > (long long)sm_count * per_sm_c) gc = (long long)sm_count * per_sm_c;
    ka<<<(int)ga, 256, kSmemA, st>>>(x, S, batch, twA, tw2d);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    kc<<<(int)gc, 256, kSmemC, st>>>(S, X, batch, twC);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

// (Nc, A1, A2): N1 = 16*A1 >= N2 = 16*A2
#define PF_T2DG_SIZES(X) X(7680, 6, 5) X(9216, 6, 6) X(12288, 8, 6) X(20480, 10, 8) X(24576, 12, 8) X(36864, 12, 12) \
                         X(40960, 16, 10) X(49152, 16, 12) X(61440, 16, 15) X(16384, 8, 8) X(32768, 16, 8) X(65536, 16, 16)

}  // namespace

bool t2dg_shape_for(int Nc, int* A1, int* A2) {
#define X(nc, a1, a2) if (Nc == nc) { *A1 = a1; *A2 = a2; return true; }
  PF_T2DG_SIZES(X)
#undef X
  return false;
}
// device tables [twA: N2][twC: N1][tw2d: Nc] for a size of the list (caller owns the allocation); nullptr on failure
cf* t2dg_make_tables_float(int Nc) {
  int a1 = 0, a2 = 0;
  if (!t2dg_shape_for(Nc, &a1, &a2)) return nullptr;
  std::vector<float> host(2 * ((size_t)16 * a1 + (size_t)16 * a2 + (size_t)Nc));
#define X(nc, a1, a2) if (Nc == nc) t2d_fill_tables<float, a1, a2>(host.data());
  PF_T2DG_SIZES(X)
#undef X
  void* d = nullptr;
  if (cudaMalloc(&d, host.size() * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(d, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tiled2d general: table allocation", cudaGetLastError());
    if (d) cudaFree(d);
    return nullptr;
  }
  return reinterpret_cast<cf*>(d);
}
int t2dg_launch_float(int Nc, int sign, const cf* x, cf* S, cf* X, long long batch, const cf* tables, int sm_count, cudaStream_t st) {
#define X(nc, a1, a2) if (Nc == nc) return sign < 0 ? T2DGLaunch<float, a1, a2, -1>::run(x, S, X, batch, tables, sm_count, st) \
                                                    : T2DGLaunch<float, a1, a2, +1>::run(x, S, X, batch, tables, sm_count, st);
  PF_T2DG_SIZES(X)
#undef X
  set_error_msg("tiled2d general: size not instantiated");
  return (int)cudaErrorInvalidValue;
}

// ---- double precision: the power-of-two shapes (PFFFT_B200_TILED2D_GENERAL=1 on a double plan)
#define PF_T2DG_SIZES_D(X) X(16384, 8, 8) X(32768, 16, 8) X(65536, 16, 16)
bool t2dg_shape_for_double(int Nc, int* A1, int* A2) {
#define X(nc, a1, a2) if (Nc == nc) { *A1 = a1; *A2 = a2; return true; }
  PF_T2DG_SIZES_D(X)
#undef X
  return false;
}
cd* t2dg_make_tables_double(int Nc) {
  int a1 = 0, a2 = 0;
  if (!t2dg_shape_for_double(Nc, &a1, &a2)) return nullptr;
  std::vector<double> host(2 * ((size_t)16 * a1 + (size_t)16 * a2 + (size_t)Nc));
#define X(nc, a1, a2) if (Nc == nc) t2d_fill_tables<double, a1, a2>(host.data());
  PF_T2DG_SIZES_D(X)
#undef X
  void* d = nullptr;
  if (cudaMalloc(&d, host.size() * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(d, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("tiled2d general: table allocation", cudaGetLastError());
    if (d) cudaFree(d);
    return nullptr;
  }
  return reinterpret_cast<cd*>(d);
}
int t2dg_launch_double(int Nc, int sign, const cd* x, cd* S, cd* X, long long batch, const cd* tables, int sm_count, cudaStream_t st) {
#define X(nc, a1, a2) if (Nc == nc) return sign < 0 ? T2DGLaunch<double, a1, a2, -1>::run(x, S, X, batch, tables, sm_count, st) \
                                                    : T2DGLaunch<double, a1, a2, +1>::run(x, S, X, batch, tables, sm_count, st);
  PF_T2DG_SIZES_D(X)
#undef X
  set_error_msg("tiled2d general (double): size not instantiated");
  return (int)cudaErrorInvalidValue;
}

}  // namespace pf
