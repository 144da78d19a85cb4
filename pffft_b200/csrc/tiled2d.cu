// tiled2d.cu -- instantiations and launchers of the tiled two-dimensional large-N plan (tiled2d_kernels.cuh), float.
// Own translation unit so the C-ABI units stay small.  Default plan for 32768 and 65536 (PFFFT_B200_TILED2D=0|1 overrides).
#define PF_NO_PACKED_F32 1   // scalar fp32 arithmetic in this translation unit: measured 4 % faster with scalar arithmetic at 16384 (cluster-fused form), equal elsewhere (profiles/r02b_packed.md)
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
    static PerDeviceInt occ_a, occ_c;
    const int dev = current_device();
    const int per_sm_a = occ_a.get(dev, [&]() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka, G::TA, kSmemA); return n < 1 ? 1 : n; });
    const int per_sm_c = occ_c.get(dev, [&]() { int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kc, G::TC, kSmemC); return n < 1 ? 1 : n; });
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

// cluster-fused form (one HBM round trip); NOT YET RUN ON HARDWARE, reached only with PFFFT_B200_TILED2D=2
template <int A1, int A2, int CL, int SIGN> struct T2DClusterLaunch {
  using G = T2D<A1, A2>;
  using K = T2DC<A1, A2, CL>;
  static constexpr size_t kSmem = K::kSmem * sizeof(cf);
  static constexpr int kBySmem = (int)((227 * 1024) / (kSmem + 1024));
  static constexpr int kByRegs = 768 / K::NT;                       // 80 registers per thread
  static constexpr int MINB = kBySmem < 1 ? 1 : (kBySmem < kByRegs ? kBySmem : kByRegs);
  static auto kernel() { return k_t2d_cluster<float, A1, A2, CL, SIGN, MINB>; }
  static int configure() {                                          // function attributes: once per device
    PF_CUDA_OK(cudaFuncSetAttribute(kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    if (CL > 8) PF_CUDA_OK(cudaFuncSetAttribute(kernel(), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    return 0;
  }
  static int prepare(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int nclusters, cudaStream_t st) {
    static PerDeviceInt configured;
    int cfg_rc = 0;
    configured.get(current_device(), [&]() -> int {
      cfg_rc = configure();
      return cfg_rc ? -1 : 1;
    });
    if (cfg_rc) return cfg_rc;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)(nclusters * CL), 1, 1);
    cfg->blockDim = dim3(K::NT, 1, 1);
    cfg->dynamicSmemBytes = kSmem;
    cfg->stream = st;
    cfg->attrs = attr; cfg->numAttrs = 1;
    return 0;
  }
  static int max_active() {
    static PerDeviceInt cached;
    return cached.get(current_device(), [&]() -> int {
      cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
      if (prepare(&cfg, attr, 1, nullptr)) return 0;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel(), &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      return n;
    });
  }
  static int run(const cf* x, cf* X, long long batch, const cf* tables, cudaStream_t st) {
    const int cap = max_active();
    if (cap <= 0) { set_error_msg("tiled2d cluster kernel: cluster shape not schedulable on this device"); return (int)cudaErrorInvalidConfiguration; }
    const int ncl = (int)(batch < (long long)cap ? batch : (long long)cap);
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    { const int rc = prepare(&cfg, attr, ncl, st); if (rc) return rc; }
    const cf* twA = tables;
    const cf* twC = twA + G::N2;
    const cf* tw2d = twC + G::N1;
    PF_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel(), x, X, batch, twA, twC, tw2d));
    count_launch();
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

namespace pf {
// cluster-fused form: 8-CTA clusters for every size
int t2d_cluster_max_active_float(int Nc) {
  switch (Nc) {
    case 16384: return T2DClusterLaunch<8, 8, 8, -1>::max_active();
    case 32768: return T2DClusterLaunch<16, 8, 8, -1>::max_active();
    case 65536: return T2DClusterLaunch<16, 16, 8, -1>::max_active();
  }
  return 0;
}
int t2d_cluster_launch_float(int Nc, int sign, const cf* x, cf* X, long long batch, const cf* tables, cudaStream_t st) {
#define PF_T2C(nc, a1, a2) if (Nc == nc) return sign < 0 ? T2DClusterLaunch<a1, a2, 8, -1>::run(x, X, batch, tables, st) \
                                                         : T2DClusterLaunch<a1, a2, 8, +1>::run(x, X, batch, tables, st);
  PF_T2C(16384, 8, 8) PF_T2C(32768, 16, 8) PF_T2C(65536, 16, 16)
#undef PF_T2C
  set_error_msg("tiled2d cluster: size not instantiated");
  return (int)cudaErrorInvalidValue;
}
}  // namespace pf
