// radix_impl.cuh -- launcher template shared by radix_a.cu / radix_b.cu (the instantiations are split over two translation
// units so they compile in parallel)
#pragma once
#include <cuda_runtime.h>
#include "internal_api.h"
#include "radix.h"
#include "radix_kernels.cuh"

namespace pf {

template <typename T, int R1, int R2, int R3, int LM, int SM, int SIGN, int TPC, int MINB>
int radix_launch_one(const T* in, T* out, long long batch, const cpx<T>* tw, const cpx<T>* twr, int device, int sm_count, cudaStream_t st) {
  using S = RadixShape<R1, R2, R3>;
  auto kern = k_cta_radix<T, R1, R2, R3, LM, SM, SIGN, TPC, MINB>;
  constexpr size_t smem = (size_t)TPC * S::NCP * sizeof(cpx<T>);
  static PerDeviceInt occ;
  int arc = 0;
  const int per_sm = occ.get(device, [&]() -> int {
    if (smem > 48 * 1024) arc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (arc) return -1;
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, TPC * S::TT, smem);
    return n < 1 ? 1 : n;
  });
  if (arc) { set_error("radix kernel: cudaFuncSetAttribute", (cudaError_t)arc); return arc; }
  long long ctas = (batch + TPC - 1) / TPC;
  const long long cap = (long long)sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, TPC * S::TT, smem, st>>>(in, out, batch, tw, twr);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}

// the eight (load, store, direction) combinations of the API for one core
template <typename T, int R1, int R2, int R3, int TPC, int MINB>
int radix_launch_modes(int lm, int sm, int sign, const T* in, T* out, long long batch, const cpx<T>* tw, const cpx<T>* twr,
                       int device, int sm_count, cudaStream_t st) {
#define PF_RX(L, S_, SG) if (lm == L && sm == S_ && sign == SG) return radix_launch_one<T, R1, R2, R3, L, S_, SG, TPC, MINB>(in, out, batch, tw, twr, device, sm_count, st);
  PF_RX(L_C_ORD, S_C_ORD, -1) PF_RX(L_C_ORD, S_C_ORD, +1) PF_RX(L_C_ORD, S_C_Z, -1) PF_RX(L_C_Z, S_C_ORD, +1)
  PF_RX(L_R_TIME, S_R_ORD, -1) PF_RX(L_R_TIME, S_R_Z, -1) PF_RX(L_R_ORD, S_R_TIME, +1) PF_RX(L_R_Z, S_R_TIME, +1)
#undef PF_RX
  return -1;
}

}  // namespace pf
