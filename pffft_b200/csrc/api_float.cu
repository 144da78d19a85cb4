// api_float.cu -- pffft_* : the single-precision C-ABI (ref include/pffft/pffft.h:124-250) and the
// selection of the size-tuned kernels.
#define PF_NO_PACKED_F32 1   // scalar fp32 arithmetic in this translation unit: the 16x16xC CTA kernels measure 1-10 % faster with scalar arithmetic (C3 backward 0.75 against 0.68) (profiles/r02b_packed.md)
#include "../../include/pffft/pffft_b200.h"
#include "api_impl.cuh"
#include "fast.h"
#include "cta_hooks.cuh"
#include "radix.h"

namespace pf {

// ---- tuned-kernel hooks for float -------------------------------------------------------------
// variant ids for c2c N=1024 (PFFFT_B200_C1024 environment variable, read when the plan is built;
// a tuning knob for profiling runs, the default is the measured best)
enum { V_LDG_4x4 = 0, V_LDG_8x2 = 1, V_BULK_8 = 2, V_BULK_12 = 3, V_BULK_4x3 = 4, V_COUNT = 5 };
static const char* kVariantName[V_COUNT] = {"c1024_warp_ldg_4w", "c1024_warp_ldg_8w", "c1024_warp_bulk_8w",
                                            "c1024_warp_bulk_12w", "c1024_warp_bulk_4w"};

// ---- the warp-per-transform kernels live in fast_pk.cu (packed f32x2 arithmetic) and fast_sc.cu (scalar arithmetic);
// `fast_packed` (fast.h) says which build serves a combination
static FastCtx fast_ctx(const Setup<float>* s) { return FastCtx{s->device, s->sm_count, s->tw_fast, s->twr}; }
template <int SIGN, bool ZIN, bool ZOUT>
static int run_c1024(Setup<float>* s, const float* in, float* out, long long batch, cudaStream_t st) {
  return fast_packed(FK_C1024, 0, ZIN, ZOUT, false) ? fast_c1024_pk(fast_ctx(s), s->fast_variant, SIGN, ZIN, ZOUT, in, out, batch, st)
                                                    : fast_c1024_sc(fast_ctx(s), s->fast_variant, SIGN, ZIN, ZOUT, in, out, batch, st);
}
template <int SIGN, bool ZIN, bool ZOUT>
static int run_wsmall(Setup<float>* s, int R2, const float* in, float* out, long long batch, cudaStream_t st) {
  return fast_packed(FK_WSMALL, R2, ZIN, ZOUT, false) ? fast_wsmall_pk(fast_ctx(s), R2, SIGN, ZIN, ZOUT, in, out, batch, st)
                                                      : fast_wsmall_sc(fast_ctx(s), R2, SIGN, ZIN, ZOUT, in, out, batch, st);
}
template <int SIGN, bool ZIN, bool ZOUT, bool REAL>
static int run_wmixed(Setup<float>* s, int R2, const float* in, float* out, long long batch, cudaStream_t st, int grp = 1) {
  return fast_packed(FK_WMIXED, R2, ZIN, ZOUT, REAL) ? fast_wmixed_pk(fast_ctx(s), R2, SIGN, ZIN, ZOUT, REAL, in, out, batch, st, grp)
                                                     : fast_wmixed_sc(fast_ctx(s), R2, SIGN, ZIN, ZOUT, REAL, in, out, batch, st, grp);
}
static int wsmall_R2_for(int N, int transform) {
  if (transform != XF_COMPLEX) return 0;
  return N == 32 ? 1 : N == 64 ? 2 : N == 128 ? 4 : N == 256 ? 8 : 0;
}

// complex N = 32*R2 (non-pow2 R2) and real N = 64*R2 (any supported R2)
static int wmixed_R2_for(int N, int transform) {
  const int Nc = transform == XF_REAL ? N / 2 : N;
  if (Nc % 32) return 0;
  const int r = Nc / 32;
  switch (r) { case 3: case 5: case 6: case 9: case 10: case 12: case 15: case 18: case 20: case 24: case 25: case 27: case 30: return r; }
  if (transform == XF_REAL && (r == 1 || r == 2 || r == 4 || r == 8)) return r;
  return 0;
}

// ---- rows of two-pass (split) float plans: CTA kernel sizes or complex warp-kernel sizes 32*R2
static int wmixed_complex_R2(int n) { return wmixed_R2_for(n, XF_COMPLEX); }
static bool is_float_row_size(int n) { return cta_C_for(n) != 0 || wmixed_complex_R2(n) != 0; }
template <int SIGN>
static int split_rows_float(Setup<float>* s, const cf* src, cf* rows, long long batch, cudaStream_t st) {
  if (cta_C_for(s->split_N2)) return split_rows_cta<float, SIGN>(s, src, rows, batch, st);
  return run_wmixed<SIGN, false, false, false>(s, wmixed_complex_R2(s->split_N2), reinterpret_cast<const float*>(src),
                                               reinterpret_cast<float*>(rows), batch * s->split_R, st, s->split_R);
}
static bool float_split_for(int N, int transform, int* R, int* N2) {
  const int Nc = transform == XF_REAL ? N / 2 : N;
  if (!getenv("PFFFT_B200_NO_FUSED_SPLIT") && split_choose_fused<float>(Nc, R, N2)) return true;   // one-kernel plan first
  return split_choose(Nc, is_float_row_size, R, N2);
}

// ---- compile-time-radix CTA kernels (radix_kernels.cuh): cores that are neither 32*R2 nor 256*C.  PFFFT_B200_RADIX=0 off.
static bool radix_wanted(int Nc) {
  static const int mode = getenv("PFFFT_B200_RADIX") ? atoi(getenv("PFFFT_B200_RADIX")) : 1;   // default on (profiles/r02_radix.md)
  // an explicit request for another plan of the same core wins over the default (7680, 9216 have a tiled plan and a one-CTA split plan)
  if (const char* e = getenv("PFFFT_B200_TILED2D_GENERAL")) { int a1, a2; if (atoi(e) != 0 && t2dg_shape_for(Nc, &a1, &a2)) return false; }
  return mode != 0 && radix_core_supported(Nc, nullptr);
}

template <> struct FastHooks<float> {
  static bool is_warp1024(int N, int transform) { return transform == XF_COMPLEX && N == 1024; }
  static size_t extra_table_cpx(int N, int transform) {
    if (ts_wanted(transform == XF_REAL ? N / 2 : N, false) || radix_wanted(transform == XF_REAL ? N / 2 : N)) return 0;
    if (is_warp1024(N, transform)) return 1024;
    if (wsmall_R2_for(N, transform) || wmixed_R2_for(N, transform)) return (size_t)(transform == XF_REAL ? N / 2 : N);
    { const int Nc = transform == XF_REAL ? N / 2 : N; int R = 0, N2 = 0;
      if (!cta_C_for(Nc) && float_split_for(N, transform, &R, &N2)) return cta_C_for(N2) ? split_table_cpx(Nc, N2) + t2d_table_cpx(Nc) : (size_t)N2; }
    return CtaOnlyHooks<float>::extra_table_cpx(N, transform);
  }
  static void fill_extra_table(int N, int transform, float* dst) {
    if (ts_wanted(transform == XF_REAL ? N / 2 : N, false) || radix_wanted(transform == XF_REAL ? N / 2 : N)) return;
    if (is_warp1024(N, transform)) {                        // tw[k2*32 + n1] = exp(-2 pi i n1 k2 / 1024)
      for (int k2 = 0; k2 < 32; ++k2)
        for (int n1 = 0; n1 < 32; ++n1) {
          long double c, sn;
          pfplan::unit_root((long long)n1 * k2, 1024, &c, &sn);
          dst[2 * (k2 * 32 + n1)] = (float)c; dst[2 * (k2 * 32 + n1) + 1] = (float)sn;
        }
      return;
    }
    if (const int R2 = wsmall_R2_for(N, transform) ? wsmall_R2_for(N, transform) : wmixed_R2_for(N, transform)) {   // tw[k2*32 + l] = exp(-2 pi i l k2 / Nc)
      const int Nc = transform == XF_REAL ? N / 2 : N;
      for (int k2 = 0; k2 < R2; ++k2)
        for (int l = 0; l < 32; ++l) {
          long double c, sn;
          pfplan::unit_root((long long)l * k2, Nc, &c, &sn);
          dst[2 * (k2 * 32 + l)] = (float)c; dst[2 * (k2 * 32 + l) + 1] = (float)sn;
        }
      return;
    }
    { const int Nc = transform == XF_REAL ? N / 2 : N; int R = 0, N2 = 0;
      if (!cta_C_for(Nc) && float_split_for(N, transform, &R, &N2)) {
        if (cta_C_for(N2)) {
          split_fill_tables<float>(Nc, N2, dst);
          if (t2d_table_cpx(Nc)) t2d_fill_tables_float(Nc, dst + 2 * split_table_cpx(Nc, N2));
          return;
        }
        const int R2 = wmixed_complex_R2(N2);                 // warp rows: tw[k2*32 + l] = exp(-2 pi i l k2 / N2)
        for (int k2 = 0; k2 < R2; ++k2)
          for (int l = 0; l < 32; ++l) {
            long double c, sn;
            pfplan::unit_root((long long)l * k2, N2, &c, &sn);
            dst[2 * (k2 * 32 + l)] = (float)c; dst[2 * (k2 * 32 + l) + 1] = (float)sn;
          }
        return;
      } }
    CtaOnlyHooks<float>::fill_extra_table(N, transform, dst);
  }
  static bool plan(Setup<float>* s) {
    if (ts_wanted(s->Nc, false)) return ts_plan<float>(s);
    if (radix_wanted(s->Nc)) {
      const char* nm = "";
      radix_core_supported(s->Nc, &nm);
      s->fast_variant = 600; s->kernel_name = nm;
      return true;
    }
    if (is_warp1024(s->N, s->transform)) {
      int v = V_LDG_4x4;
      if (const char* e = getenv("PFFFT_B200_C1024")) { v = atoi(e); if (v < 0 || v >= V_COUNT) v = V_LDG_4x4; }
      s->fast_variant = v;
      s->kernel_name = kVariantName[v];
      return true;
    }
    if (const int R2 = wmixed_R2_for(s->N, s->transform)) {
      if (getenv("PFFFT_B200_NO_WMIXED")) return false;
      static char names[64][20];
      const int slot = R2 + (s->transform == XF_REAL ? 32 : 0);
      snprintf(names[slot], sizeof(names[slot]), s->transform == XF_REAL ? "warp_real_32x%d" : "warp_32x%d", R2);
      s->kernel_name = names[slot];
      s->fast_variant = 400 + R2;
      return true;
    }
    if (const int R2 = wsmall_R2_for(s->N, s->transform)) {
      if (getenv("PFFFT_B200_NO_WSMALL")) return false;
      s->fast_variant = 200 + R2;
      s->kernel_name = R2 == 1 ? "warp_32x1" : R2 == 2 ? "warp_32x2" : R2 == 4 ? "warp_32x4" : "warp_32x8";
      return true;
    }
    { int R = 0, N2 = 0;
      if (!cta_C_for(s->Nc) && !getenv("PFFFT_B200_NO_SPLIT") && float_split_for(s->N, s->transform, &R, &N2)) {
        s->split_R = R; s->split_N2 = N2;
        s->split_fused = !getenv("PFFFT_B200_NO_FUSED_SPLIT") && split_fused_ok<float>(R, N2);
        s->fast_variant = 300;
        int CL = 0, Q = 1, mode = 0, a1 = 0, a2 = 0;
        if (t2dg_wanted(s->Nc, false) && t2dg_shape_for(s->Nc, &a1, &a2) && (s->d_aux_tables = t2dg_make_tables_float(s->Nc)) != nullptr) {
          s->split_fused = false;
          snprintf(s->name_buf, sizeof(s->name_buf), "tiled2dg_%dx%d", 16 * a1, 16 * a2);
        } else if (t2d_enabled(s->Nc) && cta_C_for(N2) && t2d_shape_for(s->Nc, &a1, &a2)) {
          s->split_t2d = true; s->split_fused = false;
          s->split_t2d_cluster = t2d_cluster_requested(s->Nc) && t2d_cluster_max_active_float(s->Nc) > 0;
          snprintf(s->name_buf, sizeof(s->name_buf), s->split_t2d_cluster ? "tiled2d_cluster8_%dx%d" : "tiled2d_%dx%d", 16 * a1, 16 * a2);
        } else if (cluster_choose(R, N2, &CL, &Q, &mode)) {
          s->split_cluster = CL; s->split_Q = Q; s->split_mode = mode; s->split_fused = false;
          snprintf(s->name_buf, sizeof(s->name_buf), "cluster%d_%dx%d%s", CL, R, N2, mode == 1 ? "_dsmem_rows" : "");
        } else
        snprintf(s->name_buf, sizeof(s->name_buf), s->split_fused ? "cta_split_%dx%d" : "split_%dx%d", R, N2);
        s->kernel_name = s->name_buf;
        return true;
      } }
    return CtaOnlyHooks<float>::plan(s);
  }
  static int run(Setup<float>* s, const float* in, float* out, long long batch, int direction, int ordered, cudaStream_t st,
                 const XformOpts& o) {
    if (s->fast_variant == 500) return ts_dispatch<float>(s, in, out, batch, direction, ordered, st, o);
    if (s->fast_variant == 600) {                           // compile-time-radix CTA kernels: dense aligned batches
      const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
      if (!plain || !vec_aligned<float>(in) || !vec_aligned<float>(out)) return -1;
      int lm = 0, sm = 0;
      ts_modes(s->transform, direction, ordered, &lm, &sm);
      return radix_launch_float(s->Nc, lm, sm, direction == DIR_FORWARD ? -1 : +1, in, out, batch, s->tw, s->twr, s->device, s->sm_count, st);
    }
    if (s->fast_variant < 100) {                            // warp-per-transform N=1024 complex: contiguous batches only
      const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
      if (!plain) return -1;
      if (direction == DIR_FORWARD) return ordered ? run_c1024<-1, false, false>(s, in, out, batch, st)
                                                   : run_c1024<-1, false, true>(s, in, out, batch, st);
      return ordered ? run_c1024<+1, false, false>(s, in, out, batch, st)
                     : run_c1024<+1, true, false>(s, in, out, batch, st);
    }
    if (s->fast_variant >= 400) {                           // non-power-of-two warp kernels: contiguous batches only
      const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
      if (!plain) return -1;
      const int R2 = s->fast_variant - 400;
      if (s->transform == XF_REAL) {
        if (direction == DIR_FORWARD) return ordered ? run_wmixed<-1, false, false, true>(s, R2, in, out, batch, st)
                                                     : run_wmixed<-1, false, true, true>(s, R2, in, out, batch, st);
        return ordered ? run_wmixed<+1, false, false, true>(s, R2, in, out, batch, st)
                       : run_wmixed<+1, true, false, true>(s, R2, in, out, batch, st);
      }
      if (direction == DIR_FORWARD) return ordered ? run_wmixed<-1, false, false, false>(s, R2, in, out, batch, st)
                                                   : run_wmixed<-1, false, true, false>(s, R2, in, out, batch, st);
      return ordered ? run_wmixed<+1, false, false, false>(s, R2, in, out, batch, st)
                     : run_wmixed<+1, true, false, false>(s, R2, in, out, batch, st);
    }
    if (s->fast_variant >= 200 && s->fast_variant < 300) {  // small warp kernels: contiguous batches only
      const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
      if (!plain) return -1;
      const int R2 = s->fast_variant - 200;
      if (direction == DIR_FORWARD) return ordered ? run_wsmall<-1, false, false>(s, R2, in, out, batch, st)
                                                   : run_wsmall<-1, false, true>(s, R2, in, out, batch, st);
      return ordered ? run_wsmall<+1, false, false>(s, R2, in, out, batch, st)
                     : run_wsmall<+1, true, false>(s, R2, in, out, batch, st);
    }
    if (s->fast_variant >= 300 && s->fast_variant < 400) {
      const XformParams<float> p = make_params(s, in, out, batch, o);
      return run_split<float>(s, split_rows_float<-1>, split_rows_float<+1>, p, direction, ordered, st);
    }
    return CtaOnlyHooks<float>::run(s, in, out, batch, direction, ordered, st, o);
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
  if (s->kind == KK_FAST && s->fast_variant >= 100 && s->fast_variant < 200) {
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
