// cluster.cu -- instantiations and launchers of the thread-block-cluster kernels (cluster_kernels.cuh): float complex
// cores 8192..65536 = (CL*Q) x 4096 in one HBM round trip.  Own translation unit so the C-ABI units stay small.
#define PF_NO_PACKED_F32 1   // scalar fp32 arithmetic in this translation unit: measured 4 % faster with scalar arithmetic (16384: 0.49 against 0.47) (profiles/r02b_packed.md)
#include <cuda_runtime.h>
#include <stdlib.h>
#include "internal_api.h"
#include "cluster_kernels.cuh"

namespace pf {
namespace {

// MODE 0: strided rows staged by cp.async; 1: rows distributed through DSMEM
template <int C, int CL, int Q, int SIGN, int MODE> struct ClusterLaunch {
  using G = KCL<C, CL, Q>;
  static constexpr size_t kSmem = (size_t)(1 + Q) * G::N2 * sizeof(cpx<float>);
  // CTAs per SM the register budget is sized for: what the shared memory admits (227 KB per SM), at most 1024 threads
  static constexpr int kBySmem = (int)((227 * 1024) / (kSmem + 1024));
  static constexpr int kByThreads = 1024 / (16 * C);
  static constexpr int MINB = kBySmem < 1 ? 1 : (kBySmem < kByThreads ? kBySmem : kByThreads);
  static auto kernel() { return k_cluster_fft<float, C, CL, Q, SIGN, MODE, MINB>; }

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
    cfg->blockDim = dim3(16 * C, 1, 1);
    cfg->dynamicSmemBytes = kSmem;
    cfg->stream = st;
    cfg->attrs = attr; cfg->numAttrs = 1;
    return 0;
  }
  // co-resident clusters of this kernel on the current device (0: the device cannot schedule the cluster shape)
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
  static int launch(const cf* src, cf* dst, long long batch, const cf* tw1, const cf* tw2, const cf* twP, cudaStream_t st) {
    const int cap = max_active();
    if (cap <= 0) { set_error_msg("cluster kernel: cluster shape not schedulable on this device"); return (int)cudaErrorInvalidConfiguration; }
    const int ncl = (int)(batch < (long long)cap ? batch : (long long)cap);
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    { const int rc = prepare(&cfg, attr, ncl, st); if (rc) return rc; }
    const float* in = reinterpret_cast<const float*>(src);
    float* out = reinterpret_cast<float*>(dst);
    PF_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel(), in, out, batch, tw1, tw2, twP));
    count_launch();
    return 0;
  }
};

// (CL, Q, mode) shapes that exist; every one is built for both directions
#define PF_CLUSTER_SHAPES(X) X(2, 1, 0) X(2, 1, 1) X(4, 1, 0) X(4, 1, 1) X(8, 1, 0) X(8, 1, 1) X(8, 2, 0) X(16, 1, 0) X(16, 1, 1) \
                             X(4, 2, 0) X(4, 4, 0)

}  // namespace

bool cluster_shape_exists(int CL, int Q, int mode) {
#define X(cl, q, sc) if (CL == cl && Q == q && mode == sc) return true;
  PF_CLUSTER_SHAPES(X)
#undef X
  return false;
}
int cluster_max_active_float(int CL, int Q, int mode) {
#define X(cl, q, sc) if (CL == cl && Q == q && mode == sc) return ClusterLaunch<16, cl, q, -1, sc>::max_active();
  PF_CLUSTER_SHAPES(X)
#undef X
  return 0;
}
int cluster_launch_float(int CL, int Q, int mode, int sign, const cf* src, cf* dst, long long batch,
                         const cf* tw1, const cf* tw2, const cf* twP, cudaStream_t st) {
#define X(cl, q, sc)                                                                                         \
  if (CL == cl && Q == q && mode == sc)                                                                   \
    return sign < 0 ? ClusterLaunch<16, cl, q, -1, sc>::launch(src, dst, batch, tw1, tw2, twP, st)           \
                    : ClusterLaunch<16, cl, q, +1, sc>::launch(src, dst, batch, tw1, tw2, twP, st);
  PF_CLUSTER_SHAPES(X)
#undef X
  set_error_msg("cluster kernel: shape not instantiated");
  return (int)cudaErrorInvalidValue;
}

}  // namespace pf
