// api_float.cu -- pffft_* : the single-precision C-ABI (ref include/pffft/pffft.h:124-250) and the
// selection of the size-tuned kernels.
#include "../../include/pffft/pffft_b200.h"
#include "api_impl.cuh"
#include "fast_kernels.cuh"
#include "cta_kernels.cuh"

namespace pf {

// ---- tuned-kernel hooks for float -------------------------------------------------------------
// variant ids for c2c N=1024 (PFFFT_B200_C1024 environment variable, read when the plan is built;
// a tuning knob for profiling runs, the default is the measured best)
enum { V_LDG_4x4 = 0, V_LDG_8x2 = 1, V_BULK_8 = 2, V_BULK_12 = 3, V_BULK_4x3 = 4, V_COUNT = 5 };
static const char* kVariantName[V_COUNT] = {"c1024_warp_ldg_4w", "c1024_warp_ldg_8w", "c1024_warp_bulk_8w",
                                            "c1024_warp_bulk_12w", "c1024_warp_bulk_4w"};

template <int SIGN, int WARPS, int MINB, bool ZIN, bool ZOUT>
static int launch_ldg(Setup<float>* s, const float* in, float* out, long long batch, cudaStream_t st) {
  auto kern = k_c1024_ldg<SIGN, WARPS, MINB, ZIN, ZOUT>;
  const size_t smem = (1024 + (size_t)WARPS * kW1024Tile) * sizeof(cf);
  static thread_local bool attr = false;
  if (!attr) { PF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  long long ctas = (batch + WARPS - 1) / WARPS;
  const long long cap = (long long)s->sm_count * MINB;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, s->tw_fast);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <int SIGN, int WARPS, int MINB, bool ZOUT>
static int launch_bulk(Setup<float>* s, const float* in, float* out, long long batch, cudaStream_t st) {
  auto kern = k_c1024_bulk<SIGN, WARPS, MINB, ZOUT>;
  const size_t smem = (1024 + (size_t)WARPS * 2 * kW1024Tile) * sizeof(cf) + (size_t)WARPS * 2 * sizeof(uint64_t);
  static thread_local bool attr = false;
  if (!attr) { PF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
  long long ctas = (batch + WARPS - 1) / WARPS;
  const long long cap = (long long)s->sm_count * MINB;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, WARPS * 32, smem, st>>>(reinterpret_cast<const cf*>(in), reinterpret_cast<cf*>(out), batch, s->tw_fast);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int SIGN, bool ZIN, bool ZOUT>
static int run_c1024(Setup<float>* s, const float* in, float* out, long long batch, cudaStream_t st) {
  switch (s->fast_variant) {
    case V_LDG_8x2: return launch_ldg<SIGN, 8, 2, ZIN, ZOUT>(s, in, out, batch, st);
    case V_BULK_8:  if (!ZIN) return launch_bulk<SIGN, 8, 1, ZOUT>(s, in, out, batch, st); break;
    case V_BULK_12: if (!ZIN) return launch_bulk<SIGN, 12, 1, ZOUT>(s, in, out, batch, st); break;
    case V_BULK_4x3: if (!ZIN) return launch_bulk<SIGN, 4, 3, ZOUT>(s, in, out, batch, st); break;
    default: break;
  }
  return launch_ldg<SIGN, 4, 4, ZIN, ZOUT>(s, in, out, batch, st);
}

#ifndef PF_CTA_TPSM
#define PF_CTA_TPSM 1024
#endif
// ---- CTA-per-transform kernels (cta_kernels.cuh): complex cores of 512 / 1024 / 2048 / 4096 points
template <int C, int LM, int SM, int SIGN, bool STAGED>
static int launch_cta_v(Setup<float>* s, const XformParams<float>& p, cudaStream_t st) {
  constexpr int MINB = PF_CTA_TPSM / (16 * C);              // threads per SM the register budget is sized for
  auto kern = k_cta_fft<C, LM, SM, SIGN, MINB, STAGED>;
  const size_t smem = (size_t)K2<C>::NC * sizeof(cf) * (STAGED ? 2 : 1) + (STAGED ? 16 : 0);
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    if (smem > 48 * 1024) PF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 16 * C, smem);
    if (per_sm < 1) per_sm = 1;
  }
  long long ctas = p.batch;
  const long long cap = (long long)s->sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  const cf* tw1 = s->tw_fast;
  const cf* tw2 = s->tw_fast + K2<C>::NC;
  kern<<<(int)ctas, 16 * C, smem, st>>>(p, tw1, tw2);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
// STAGED (TMA-fed) variant when the call is a plain contiguous, 16-byte aligned batch in canonical input order
// measured slower than register-fed loads at 1024 threads/SM (C3: 0.70 vs 0.79 of HBM peak): opt-in only
static bool g_cta_stage = getenv("PFFFT_B200_CTA_STAGE") ? atoi(getenv("PFFFT_B200_CTA_STAGE")) != 0 : false;
template <int C, int LM, int SM, int SIGN>
static int launch_cta(Setup<float>* s, const XformParams<float>& p, cudaStream_t st) {
  if constexpr (LM == L_C_ORD || LM == L_R_TIME) {
    const bool contiguous = p.in_limit < 0 && p.in_stride == (long long)s->per() && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
    if (g_cta_stage && contiguous) return launch_cta_v<C, LM, SM, SIGN, true>(s, p, st);
  }
  return launch_cta_v<C, LM, SM, SIGN, false>(s, p, st);
}
template <int C>
static int run_cta(Setup<float>* s, const XformParams<float>& p, int direction, int ordered, cudaStream_t st) {
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? launch_cta<C, L_C_ORD, S_C_ORD, -1>(s, p, st) : launch_cta<C, L_C_ORD, S_C_Z, -1>(s, p, st);
    return ordered ? launch_cta<C, L_C_ORD, S_C_ORD, +1>(s, p, st) : launch_cta<C, L_C_Z, S_C_ORD, +1>(s, p, st);
  }
  if (fwd) return ordered ? launch_cta<C, L_R_TIME, S_R_ORD, -1>(s, p, st) : launch_cta<C, L_R_TIME, S_R_Z, -1>(s, p, st);
  return ordered ? launch_cta<C, L_R_ORD, S_R_TIME, +1>(s, p, st) : launch_cta<C, L_R_Z, S_R_TIME, +1>(s, p, st);
}
static int cta_C_for(int Nc) { return Nc == 512 ? 2 : Nc == 1024 ? 4 : Nc == 2048 ? 8 : Nc == 4096 ? 16 : 0; }

template <> struct FastHooks<float> {
  static bool is_warp1024(int N, int transform) { return transform == XF_COMPLEX && N == 1024; }
  static size_t extra_table_cpx(int N, int transform) {
    if (is_warp1024(N, transform)) return 1024;
    const int Nc = transform == XF_REAL ? N / 2 : N;
    const int C = cta_C_for(Nc);
    return C ? (size_t)Nc + 16 * (size_t)C : 0;
  }
  static void fill_extra_table(int N, int transform, float* dst) {
    if (is_warp1024(N, transform)) {                        // tw[k2*32 + n1] = exp(-2 pi i n1 k2 / 1024)
      for (int k2 = 0; k2 < 32; ++k2)
        for (int n1 = 0; n1 < 32; ++n1) {
          long double c, sn;
          pfplan::unit_root((long long)n1 * k2, 1024, &c, &sn);
          dst[2 * (k2 * 32 + n1)] = (float)c; dst[2 * (k2 * 32 + n1) + 1] = (float)sn;
        }
      return;
    }
    const int Nc = transform == XF_REAL ? N / 2 : N;
    const int C = cta_C_for(Nc);
    if (!C) return;
    const int BC = 16 * C;
    for (int ka = 0; ka < 16; ++ka)                         // tw1[ka*BC + m] = exp(-2 pi i m ka / Nc)
      for (int m = 0; m < BC; ++m) {
        long double c, sn;
        pfplan::unit_root((long long)m * ka, Nc, &c, &sn);
        dst[2 * (ka * BC + m)] = (float)c; dst[2 * (ka * BC + m) + 1] = (float)sn;
      }
    float* t2 = dst + 2 * (size_t)Nc;
    for (int kb = 0; kb < 16; ++kb)                         // tw2[kb*C + nc] = exp(-2 pi i nc kb / BC)
      for (int nc = 0; nc < C; ++nc) {
        long double c, sn;
        pfplan::unit_root((long long)nc * kb, BC, &c, &sn);
        t2[2 * (kb * C + nc)] = (float)c; t2[2 * (kb * C + nc) + 1] = (float)sn;
      }
  }
  static bool plan(Setup<float>* s) {
    if (is_warp1024(s->N, s->transform)) {
      int v = V_LDG_4x4;
      if (const char* e = getenv("PFFFT_B200_C1024")) { v = atoi(e); if (v < 0 || v >= V_COUNT) v = V_LDG_4x4; }
      s->fast_variant = v;
      s->kernel_name = kVariantName[v];
      return true;
    }
    const int C = cta_C_for(s->Nc);
    if (!C || getenv("PFFFT_B200_NO_CTA")) return false;
    s->fast_variant = 100 + C;
    s->kernel_name = C == 2 ? "cta_16x16x2" : C == 4 ? "cta_16x16x4" : C == 8 ? "cta_16x16x8" : "cta_16x16x16";
    return true;
  }
  static int run(Setup<float>* s, const float* in, float* out, long long batch, int direction, int ordered, cudaStream_t st,
                 const XformOpts& o) {
    if (s->fast_variant < 100) {                            // warp-per-transform N=1024 complex: contiguous batches only
      const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
      if (!plain) return -1;
      if (direction == DIR_FORWARD) return ordered ? run_c1024<-1, false, false>(s, in, out, batch, st)
                                                   : run_c1024<-1, false, true>(s, in, out, batch, st);
      return ordered ? run_c1024<+1, false, false>(s, in, out, batch, st)
                     : run_c1024<+1, true, false>(s, in, out, batch, st);
    }
    const XformParams<float> p = make_params(s, in, out, batch, o);
    switch (s->fast_variant - 100) {
      case 2: return run_cta<2>(s, p, direction, ordered, st);
      case 4: return run_cta<4>(s, p, direction, ordered, st);
      case 8: return run_cta<8>(s, p, direction, ordered, st);
      default: return run_cta<16>(s, p, direction, ordered, st);
    }
  }
};

}  // namespace pf

PF_API(pffft_, pffftb_, PFFFT_Setup, float, pf::FastHooks<float>, floats_per_transform)

extern "C" PFFFT_EXPORT int pffftb_setup_device(const PFFFT_Setup* s) { return s ? s->device : -1; }

// ---- internal C++ entry points used by fastconv.cu (see internal_api.h)
#include "internal_api.h"
namespace pf {
int float_transform_device(PFFFT_Setup* s, const float* in, float* out, long long batch, int direction, int ordered,
                           cudaStream_t st, const XformOpts& o) {
  return engine_transform_device<float, FastHooks<float>>(s, in, out, batch, direction, ordered, st, o);
}
FloatPlanTables float_plan_tables(PFFFT_Setup* s) {
  FloatPlanTables t{0, s->sm_count, nullptr, nullptr, s->twr};
  if (s->kind == KK_FAST && s->fast_variant >= 100) {
    t.C = s->fast_variant - 100;
    t.tw1 = s->tw_fast;
    t.tw2 = s->tw_fast + s->Nc;
  }
  return t;
}
int float_zreorder_device(PFFFT_Setup* s, const float* in, float* out, long long batch, int direction, cudaStream_t st) {
  return engine_zreorder_device<float>(s, in, out, batch, direction, st);
}
int float_zconvolve_device(PFFFT_Setup* s, const float* a, const float* b, float* ab, float scaling, long long batch,
                           int b_shared, int accumulate, cudaStream_t st) {
  return engine_zconvolve_device<float>(s, a, b, ab, scaling, batch, b_shared, accumulate, st);
}
}  // namespace pf
