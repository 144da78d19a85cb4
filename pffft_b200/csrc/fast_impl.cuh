// fast_impl.cuh -- launchers of the warp-per-transform float kernels; included by fast_pk.cu / fast_sc.cu, each of which
// instantiates the kernels `fast_packed` assigns to it (PF_FAST_PART: 1 = packed build, 0 = scalar build).
#pragma once
#include <cuda_runtime.h>
#include "engine.cuh"
#include "fast.h"
#include "fast_kernels.cuh"

namespace pf {
namespace {

#if PF_FAST_PART
#define PF_FAST_FN(name) name##_pk
constexpr bool kPart = true;
#else
#define PF_FAST_FN(name) name##_sc
constexpr bool kPart = false;
#endif
// variant ids for c2c N=1024 (PFFFT_B200_C1024, read when the plan is built; the default is the measured best)
enum { V_LDG_4x4 = 0, V_LDG_8x2 = 1, V_BULK_8 = 2, V_BULK_12 = 3, V_BULK_4x3 = 4 };

template <int SIGN, int WARPS, int MINB, bool ZIN, bool ZOUT>
int launch_ldg(const FastCtx& c, const float* in, float* out, long long batch, cudaStream_t st) {
  auto kern = k_c1024_ldg<SIGN, WARPS, MINB, ZIN, ZOUT>;
  const size_t smem = (1024 + (size_t)WARPS * kW1024Tile) * sizeof(cf);
  static PerDeviceInt attr;
  { const int rc = ensure_dyn_smem(attr, c.device, kern, smem); if (rc) return rc; }
  long long ctas = (batch + WARPS - 1) / WARPS;
  const long long cap = (long long)c.sm_count * MINB;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, c.tw_fast);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int SIGN, int WARPS, int MINB, bool ZOUT>
int launch_bulk(const FastCtx& c, const float* in, float* out, long long batch, cudaStream_t st) {
  auto kern = k_c1024_bulk<SIGN, WARPS, MINB, ZOUT>;
  const size_t smem = (1024 + (size_t)WARPS * 2 * kW1024Tile) * sizeof(cf) + (size_t)WARPS * 2 * sizeof(uint64_t);
  static PerDeviceInt attr;
  { const int rc = ensure_dyn_smem(attr, c.device, kern, smem); if (rc) return rc; }
  long long ctas = (batch + WARPS - 1) / WARPS;
  const long long cap = (long long)c.sm_count * MINB;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, c.tw_fast);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int SIGN, bool ZIN, bool ZOUT>
int run_c1024(const FastCtx& c, int variant, const float* in, float* out, long long batch, cudaStream_t st) {
  if constexpr (fast_packed(FK_C1024, 0, ZIN, ZOUT, false) == kPart) {
    switch (variant) {
      case V_LDG_8x2: return launch_ldg<SIGN, 8, 2, ZIN, ZOUT>(c, in, out, batch, st);
      case V_BULK_8:  if (!ZIN) return launch_bulk<SIGN, 8, 1, ZOUT>(c, in, out, batch, st); break;
      case V_BULK_12: if (!ZIN) return launch_bulk<SIGN, 12, 1, ZOUT>(c, in, out, batch, st); break;
      case V_BULK_4x3: if (!ZIN) return launch_bulk<SIGN, 4, 3, ZOUT>(c, in, out, batch, st); break;
      default: break;
    }
    return launch_ldg<SIGN, 4, 4, ZIN, ZOUT>(c, in, out, batch, st);
  } else return -1;
}

// ---- small complex sizes on the warp machinery (N = 32..256)
template <int R2, int SIGN, bool ZIN, bool ZOUT>
int launch_wsmall(const FastCtx& c, const float* in, float* out, long long batch, cudaStream_t st) {
  if constexpr (fast_packed(FK_WSMALL, R2, ZIN, ZOUT, false) == kPart) {
    constexpr int WARPS = 4, MINB = 4;
    auto kern = k_warp_small<R2, SIGN, WARPS, MINB, ZIN, ZOUT>;
    const size_t smem = (32 * R2 + (size_t)WARPS * kW1024Tile) * sizeof(cf);
    const long long nchunks = (batch + (32 / R2) - 1) / (32 / R2);
    long long ctas = (nchunks + WARPS - 1) / WARPS;
    const long long cap = (long long)c.sm_count * MINB;
    if (ctas > cap) ctas = cap;
    kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, c.tw_fast);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    return 0;
  } else return -1;
}
template <int SIGN, bool ZIN, bool ZOUT>
int run_wsmall(const FastCtx& c, int R2, const float* in, float* out, long long batch, cudaStream_t st) {
  switch (R2) {
    case 1: return launch_wsmall<1, SIGN, ZIN, ZOUT>(c, in, out, batch, st);
    case 2: return launch_wsmall<2, SIGN, ZIN, ZOUT>(c, in, out, batch, st);
    case 4: return launch_wsmall<4, SIGN, ZIN, ZOUT>(c, in, out, batch, st);
    case 8: return launch_wsmall<8, SIGN, ZIN, ZOUT>(c, in, out, batch, st);
    default: return -1;
  }
}

// ---- non-power-of-two complex sizes (N = 32*R2) and real sizes N = 64*R2 on the warp machinery
template <int R2, int SIGN, bool ZIN, bool ZOUT, bool REAL>
int launch_wmixed(const FastCtx& c, const float* in, float* out, long long batch, cudaStream_t st, int grp) {
  if constexpr (fast_packed(FK_WMIXED, R2, ZIN, ZOUT, REAL) == kPart) {
    constexpr int WARPS = 4, MINB = 4;
    auto kern = k_warp_mixed<R2, SIGN, WARPS, MINB, ZIN, ZOUT, REAL>;
    const size_t smem = (32 * R2 + (size_t)WARPS * kW1024Tile) * sizeof(cf);
    const long long nchunks = (batch + (32 / R2) - 1) / (32 / R2);
    long long ctas = (nchunks + WARPS - 1) / WARPS;
    const long long cap = (long long)c.sm_count * MINB;
    if (ctas > cap) ctas = cap;
    kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, c.tw_fast, c.twr, grp);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    return 0;
  } else return -1;
}
template <int SIGN, bool ZIN, bool ZOUT, bool REAL>
int run_wmixed(const FastCtx& c, int R2, const float* in, float* out, long long batch, cudaStream_t st, int grp) {
  switch (R2) {
#define PF_WM(r) case r: return launch_wmixed<r, SIGN, ZIN, ZOUT, REAL>(c, in, out, batch, st, grp);
    PF_WM(3) PF_WM(5) PF_WM(6) PF_WM(9) PF_WM(10) PF_WM(12) PF_WM(15) PF_WM(18) PF_WM(20) PF_WM(24) PF_WM(25) PF_WM(27) PF_WM(30)
#undef PF_WM
    default: break;
  }
  if constexpr (REAL) {                                     // power-of-two packed lengths only exist as real plans here
    switch (R2) {
      case 1: return launch_wmixed<1, SIGN, ZIN, ZOUT, true>(c, in, out, batch, st, grp);
      case 2: return launch_wmixed<2, SIGN, ZIN, ZOUT, true>(c, in, out, batch, st, grp);
      case 4: return launch_wmixed<4, SIGN, ZIN, ZOUT, true>(c, in, out, batch, st, grp);
      case 8: return launch_wmixed<8, SIGN, ZIN, ZOUT, true>(c, in, out, batch, st, grp);
      default: break;
    }
  }
  return -1;
}

// (sign, zin, zout) -> template arguments: forward reads canonical input, backward writes canonical output
#define PF_FAST_MODES(CALL)                                            \
  if (sign < 0 && !zin && !zout) return CALL(-1, false, false);        \
  if (sign < 0 && !zin && zout)  return CALL(-1, false, true);         \
  if (sign > 0 && !zin && !zout) return CALL(+1, false, false);        \
  if (sign > 0 && zin && !zout)  return CALL(+1, true, false);         \
  return -1;

}  // namespace

int PF_FAST_FN(fast_c1024)(const FastCtx& c, int variant, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st) {
#define PF_CALL(S, ZI, ZO) run_c1024<S, ZI, ZO>(c, variant, in, out, batch, st)
  PF_FAST_MODES(PF_CALL)
#undef PF_CALL
}
int PF_FAST_FN(fast_wsmall)(const FastCtx& c, int R2, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st) {
#define PF_CALL(S, ZI, ZO) run_wsmall<S, ZI, ZO>(c, R2, in, out, batch, st)
  PF_FAST_MODES(PF_CALL)
#undef PF_CALL
}
int PF_FAST_FN(fast_wmixed)(const FastCtx& c, int R2, int sign, bool zin, bool zout, bool real, const float* in, float* out, long long batch, cudaStream_t st, int grp) {
  if (real) {
#define PF_CALL(S, ZI, ZO) run_wmixed<S, ZI, ZO, true>(c, R2, in, out, batch, st, grp)
    PF_FAST_MODES(PF_CALL)
#undef PF_CALL
  }
#define PF_CALL(S, ZI, ZO) run_wmixed<S, ZI, ZO, false>(c, R2, in, out, batch, st, grp)
  PF_FAST_MODES(PF_CALL)
#undef PF_CALL
}

}  // namespace pf
