// cta_hooks.cuh -- plan tables and launchers of the 16x16xC CTA kernels, shared by the float and double APIs.
#pragma once
#include "engine.cuh"
#include "cta_kernels.cuh"
#include "radix.h"

namespace pf {

inline int cta_C_for(int Nc) { return Nc == 512 ? 2 : Nc == 1024 ? 4 : Nc == 2048 ? 8 : Nc == 4096 ? 16 : 0; }
inline const char* cta_name(int C) { return C == 2 ? "cta_16x16x2" : C == 4 ? "cta_16x16x4" : C == 8 ? "cta_16x16x8" : "cta_16x16x16"; }
inline size_t cta_table_cpx(int Nc) { const int C = cta_C_for(Nc); return C ? (size_t)Nc + 16 * (size_t)C : 0; }

// [tw1: Nc][tw2: 16*C]   tw1[ka*BC + m] = exp(-2 pi i m ka / Nc),  tw2[kb*C + nc] = exp(-2 pi i nc kb / BC)
template <typename T> void cta_fill_tables(int Nc, T* dst) {
  const int C = cta_C_for(Nc);
  if (!C) return;
  const int BC = 16 * C;
  for (int ka = 0; ka < 16; ++ka)
    for (int m = 0; m < BC; ++m) {
      long double c, sn;
      pfplan::unit_root((long long)m * ka, Nc, &c, &sn);
      dst[2 * (ka * BC + m)] = (T)c; dst[2 * (ka * BC + m) + 1] = (T)sn;
    }
  T* t2 = dst + 2 * (size_t)Nc;
  for (int kb = 0; kb < 16; ++kb)
    for (int nc = 0; nc < C; ++nc) {
      long double c, sn;
      pfplan::unit_root((long long)nc * kb, BC, &c, &sn);
      t2[2 * (kb * C + nc)] = (T)c; t2[2 * (kb * C + nc) + 1] = (T)sn;
    }
}

// threads per SM the register budget is sized for: 1024 (64 regs) float, 512 (128 regs) double
template <typename T> constexpr int cta_tpsm() { return sizeof(T) == 4 ? 1024 : 512; }

template <typename T, int C, int LM, int SM, int SIGN, bool STAGED>
int launch_cta_v(Setup<T>* s, const XformParams<T>& p, cudaStream_t st) {
  constexpr int MINB = cta_tpsm<T>() / (16 * C);
  auto kern = k_cta_fft<T, C, LM, SM, SIGN, MINB, STAGED>;
  // second buffer: TMA stage, gather staging of z-domain / backward-real inputs, or the z image of a C=16 real forward
  constexpr bool kSecond = STAGED || LM == L_C_Z || LM == L_R_Z || (SM == S_R_Z && C == 16);
  const size_t smem = (size_t)K2<C>::NC * sizeof(cpx<T>) * (kSecond ? 2 : 1) + (STAGED ? 16 : 0);
  static PerDeviceInt cache;                                   // resident CTAs per SM on each device (attribute set first)
  int attr_rc = 0;
  const int per_sm = cache.get(s->device, [&]() -> int {
    if (smem > 48 * 1024) attr_rc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr_rc) return -1;
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 16 * C, smem);
    return n < 1 ? 1 : n;
  });
  if (attr_rc) { set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", (cudaError_t)attr_rc); return attr_rc; }
  long long ctas = p.batch;
  const long long cap = (long long)s->sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  const cpx<T>* tw1 = s->tw_fast;
  const cpx<T>* tw2 = s->tw_fast + K2<C>::NC;
  kern<<<(int)ctas, 16 * C, smem, st>>>(p, tw1, tw2);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
// TMA-staged inputs measured slower than register-fed loads at 1024 threads/SM (C3: 0.70 vs 0.79 of HBM peak):
// opt-in with PFFFT_B200_CTA_STAGE=1, float only, contiguous 16-byte aligned canonical input
inline bool cta_stage_enabled() {
  static const bool on = getenv("PFFFT_B200_CTA_STAGE") ? atoi(getenv("PFFFT_B200_CTA_STAGE")) != 0 : false;
  return on;
}
template <typename T, int C, int LM, int SM, int SIGN>
int launch_cta(Setup<T>* s, const XformParams<T>& p, cudaStream_t st) {
  if constexpr (sizeof(T) == 4 && (LM == L_C_ORD || LM == L_R_TIME)) {
    const bool contiguous = p.in_limit < 0 && p.in_stride == (long long)s->per() && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0;
    if (cta_stage_enabled() && contiguous) return launch_cta_v<T, C, LM, SM, SIGN, true>(s, p, st);
  }
  return launch_cta_v<T, C, LM, SM, SIGN, false>(s, p, st);
}
template <typename T, int C>
int run_cta(Setup<T>* s, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? launch_cta<T, C, L_C_ORD, S_C_ORD, -1>(s, p, st) : launch_cta<T, C, L_C_ORD, S_C_Z, -1>(s, p, st);
    return ordered ? launch_cta<T, C, L_C_ORD, S_C_ORD, +1>(s, p, st) : launch_cta<T, C, L_C_Z, S_C_ORD, +1>(s, p, st);
  }
  if (fwd) return ordered ? launch_cta<T, C, L_R_TIME, S_R_ORD, -1>(s, p, st) : launch_cta<T, C, L_R_TIME, S_R_Z, -1>(s, p, st);
  return ordered ? launch_cta<T, C, L_R_ORD, S_R_TIME, +1>(s, p, st) : launch_cta<T, C, L_R_Z, S_R_TIME, +1>(s, p, st);
}
template <typename T>
int run_cta_any(Setup<T>* s, int C, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  switch (C) {
    case 2: return run_cta<T, 2>(s, p, direction, ordered, st);
    case 4: return run_cta<T, 4>(s, p, direction, ordered, st);
    case 8: return run_cta<T, 8>(s, p, direction, ordered, st);
    default: return run_cta<T, 16>(s, p, direction, ordered, st);
  }
}


// ---------------------------------------------------------------- two-pass plans: Nc = R x N2
// Sizes without a single-kernel path (complex cores > 4096, non-power-of-two cores > 1024) are split by decimation in
// time over the first digit:
//   rows   : ONE launch of a tuned kernel over batch*R rows; row n1 of a transform = FFT_N2 of x[n1 + R*n2] (element
//            stride R; consecutive CTAs/warps take the R sub-sequences of the same transform so their strided reads of
//            the same 128-byte lines meet in L2)
//   combine: radix-R register DFTs with W_Nc^{n1 k2} twiddles finish the transform in natural order (coalesced).
// Two HBM round trips (plus one for a real pre-/post-rotation or z-domain pass): ceiling 0.5 of the roofline, against
// (stages+2) round trips of the global Stockham path it replaces.  R is any size of the register DFT library
// (2,3,4,5,6,8,9,10,12,15,16); N2 is a CTA-kernel size (512..4096) or, for float, a warp-kernel size 32*R2.
static const int kSplitRadices[] = {2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16};
template <typename T> inline bool split_fused_ok_fwd(int R, int N2);   // defined with the fused launcher below
template <typename IsRow> inline bool split_choose(int Nc, IsRow is_row_size, int* R, int* N2) {
  for (int r : kSplitRadices) if (Nc % r == 0 && is_row_size(Nc / r)) { *R = r; *N2 = Nc / r; return true; }
  return false;
}
// decomposition that the single-kernel variant can run, if any
template <typename T> inline bool split_choose_fused(int Nc, int* R, int* N2) {
  for (int r : kSplitRadices) if (Nc % r == 0 && split_fused_ok_fwd<T>(r, Nc / r)) { *R = r; *N2 = Nc / r; return true; }
  return false;
}

template <typename T, int SIGN>
int split_rows_cta(Setup<T>* s, const cpx<T>* src, cpx<T>* rows, long long batch, cudaStream_t st) {
  const int R = s->split_R, N2 = s->split_N2;
  XformParams<T> q;
  q.in = reinterpret_cast<const T*>(src); q.out = reinterpret_cast<T*>(rows);
  q.in_stride = 2LL * s->Nc; q.in_group = R; q.in_gstep = 2; q.in_estride = R;
  q.out_stride = 2LL * N2; q.in_limit = -1; q.out_count = 2 * N2;
  q.batch = batch * R; q.N = N2; q.Nc = N2; q.nfac = 0; q.tw = s->tw; q.twr = nullptr; q.magic_nc = 0;
  for (int i = 0; i < PF_MAX_FACTORS; ++i) { q.fac[i] = 1; q.magic[i] = 0; }
  switch (cta_C_for(N2)) {
    case 2: return launch_cta_v<T, 2, L_C_ORD, S_C_ORD, SIGN, false>(s, q, st);
    case 4: return launch_cta_v<T, 4, L_C_ORD, S_C_ORD, SIGN, false>(s, q, st);
    case 8: return launch_cta_v<T, 8, L_C_ORD, S_C_ORD, SIGN, false>(s, q, st);
    default: return launch_cta_v<T, 16, L_C_ORD, S_C_ORD, SIGN, false>(s, q, st);
  }
}
template <typename T, int SIGN>
int split_combine(Setup<T>* s, const cpx<T>* rows, cpx<T>* dst, long long batch, cudaStream_t st) {
  const int R = s->split_R, N2 = s->split_N2;
  const long long work = batch * N2;
  long long g = (work + 255) / 256; const long long cap = (long long)s->sm_count * 16;
  if (g > cap) g = cap; if (g < 1) g = 1;
  // combine twiddles W_Nc^{n1 k2}: row-major copy (unit-stride reads) when the plan carries one (CTA-core rows), else the
  // natural table exp(-2 pi i k / Nc) read at n1*k2
  const bool rm = cta_C_for(N2) != 0;
  const cpx<T>* tw = rm ? s->tw_fast + cta_table_cpx(N2) : s->tw;
  switch (R) {
#define PF_CMB(r) case r: if (rm) k_split_combine_any<T, r, SIGN, true><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, tw); \
                          else k_split_combine_any<T, r, SIGN, false><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, tw); break;
    PF_CMB(2) PF_CMB(3) PF_CMB(4) PF_CMB(5) PF_CMB(6) PF_CMB(8) PF_CMB(9) PF_CMB(10) PF_CMB(12) PF_CMB(15)
#undef PF_CMB
    default: if (rm) k_split_combine<T, 16, SIGN, true><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, tw);
             else k_split_combine<T, 16, SIGN, false><<<(int)g, 256, 0, st>>>(rows, dst, batch, N2, tw);
             break;
  }
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}

// (A decimation-in-frequency order -- dense radix-R pre-pass, then rows whose STORES carry the element stride R -- was
//  built and measured: 0.25 / 0.22 / 0.17 of HBM peak at 16384 / 32768 / 65536 against 0.38 / 0.34 / 0.24 for this
//  order.  Partial-sector stores cost more than the strided loads they replace; stores stay dense.)

// ---- tiled Stockham pipeline (ts_kernels.cuh / ts.cu): complex cores that factor into 2-4 radices 16*A, one persistent
// kernel, intermediates resident in L2.  PFFFT_B200_TS=0 never, =1 every factorisable core >= PFFFT_B200_TS_MIN (default
// 8192); unset: the measured default range.
inline bool ts_wanted(int Nc, bool dbl) {
  // measured defaults (profiles/r02_large_n.md): everything above 65536 (the global multi-launch path it replaces ran at
  // 0.08 of the roofline), and the double core 16384 (0.37 against 0.34 for the two-launch split plan)
  const bool other_choice = getenv("PFFFT_B200_TILED2D_GENERAL") || getenv("PFFFT_B200_TILED2D");   // explicit plan switches win
  bool want = Nc >= 131072 || (dbl && Nc == 16384 && !other_choice);
  // round 2b: float cores 15360 ... 131071 WITHOUT a tuned plan of their own (tiled 2-D: 16384, 32768, 65536; general-radix
  // tiled: 20480 ... 61440) ran as two launches through HBM (decimated rows + combine: 0.23-0.35 of the roofline) or on the
  // multi-launch global path (0.08); the pipeline measured 0.35-0.45 at the two-pass sizes (profiles/r02b_large_n.md)
  if (!other_choice && (Nc >= 15360 || Nc == 12800) && Nc < 131072) {          // (12800 = 160 x 80: split_16x800 0.27 -> 0.41)
    if (dbl) want = want || (Nc != 32768 && Nc != 65536);          // double: the general-radix tiled plan keeps those two
    else switch (Nc) {
      case 16384: case 20480: case 24576: case 32768: case 36864: case 40960: case 49152: case 61440: case 65536: break;
      default: want = true;
    }
  }
  if (const char* e = getenv("PFFFT_B200_TS")) {
    if (atoi(e) == 0) return false;
    int lo = 8192;
    if (const char* m = getenv("PFFFT_B200_TS_MIN")) lo = atoi(m);
    want = Nc >= lo;
  }
  if (!want) return false;
  int P = 0, A[4];
  return ts_factorize(Nc, &P, A);
}
// LoadMode / StoreMode of a dense call
inline void ts_modes(int transform, int direction, int ordered, int* lm, int* sm) {
  const bool fwd = direction == DIR_FORWARD;
  if (transform == XF_COMPLEX) { *lm = (fwd || ordered) ? L_C_ORD : L_C_Z; *sm = (fwd && !ordered) ? S_C_Z : S_C_ORD; }
  else if (fwd) { *lm = L_R_TIME; *sm = ordered ? S_R_ORD : S_R_Z; }
  else { *lm = ordered ? L_R_ORD : L_R_Z; *sm = S_R_TIME; }
}
template <typename T> inline bool ts_wanted_for(int Nc) { return ts_wanted(Nc, sizeof(T) == 8); }
template <typename T> inline bool ts_plan(Setup<T>* s) {
  s->ts = ts_create(s->N, s->Nc, sizeof(T) == 8, s->device, s->sm_count);
  if (!s->ts) return false;
  s->fast_variant = 500;
  s->kernel_name = ts_name(s->ts);
  return true;
}
template <typename T> inline int ts_dispatch(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered,
                                             cudaStream_t st, const XformOpts& o) {
  const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
  if (!plain || !vec_aligned<T>(in) || !vec_aligned<T>(out)) return -1;   // generic path
  int lm = 0, sm = 0;
  ts_modes(s->transform, direction, ordered, &lm, &sm);
  return ts_run<T>(s->ts, in, out, batch, direction == DIR_FORWARD ? -1 : +1, lm, sm, s->tw, s->twr, st);
}

// ---- tiled two-dimensional plan (tiled2d_kernels.cuh, instantiated in tiled2d.cu): float complex cores 16384 / 32768 / 65536
// as N1 x N2 with 128-byte runs in both passes.  Measured (profiles/r01b_large_n.md): 32768: 0.42, 65536: 0.41-0.44 of HBM
// peak against 0.34 / 0.26 for the split plan -> the default there; 16384: 0.39 against 0.41 for the 4-CTA cluster kernel
// -> only with PFFFT_B200_TILED2D=1.  PFFFT_B200_TILED2D=0 switches the plan off.
bool t2d_shape_for(int Nc, int* A1, int* A2);
size_t t2d_table_cpx(int Nc);                                   // [twA: N2][twC: N1][tw2d: Nc], 0 when the size has no tiled plan
void t2d_fill_tables_float(int Nc, float* dst);
int t2d_launch_float(int Nc, int sign, const cpx<float>* x, cpx<float>* S, cpx<float>* X, long long batch,
                     const cpx<float>* tables, int sm_count, cudaStream_t st);
// general-radix form (Nc = 256*A1*A2: 7680, 9216, 12288, 20480, 24576, 36864, 40960, 49152, 61440 ...; tiled2d_general.cu):
// verified by CPU stepping, NOT YET RUN ON HARDWARE -> only with PFFFT_B200_TILED2D_GENERAL=1.  Its tables are a separate
// allocation made when the plan is chosen, so plans without the switch are byte-for-byte what they were.
bool t2dg_shape_for(int Nc, int* A1, int* A2);
cpx<float>* t2dg_make_tables_float(int Nc);
int t2dg_launch_float(int Nc, int sign, const cpx<float>* x, cpx<float>* S, cpx<float>* X, long long batch,
                      const cpx<float>* tables, int sm_count, cudaStream_t st);
bool t2dg_shape_for_double(int Nc, int* A1, int* A2);            // double: the power-of-two shapes
cpx<double>* t2dg_make_tables_double(int Nc);
int t2dg_launch_double(int Nc, int sign, const cpx<double>* x, cpx<double>* S, cpx<double>* X, long long batch,
                       const cpx<double>* tables, int sm_count, cudaStream_t st);
// PFFFT_B200_TILED2D_GENERAL=0 never, =1 every shape that exists; unset: where it measured faster than the split / tiled
// plans on hardware (profiles/r02_large_n.md): float 20480 .. 65536 except the cores a single CTA or cluster holds; double 32768, 65536
inline bool t2dg_wanted(int Nc, bool dbl) {
  if (const char* e = getenv("PFFFT_B200_TILED2D_GENERAL")) return atoi(e) != 0;
  if (getenv("PFFFT_B200_TILED2D")) return false;                 // an explicit choice among the other tiled plans wins
  if (dbl) return Nc == 32768 || Nc == 65536;
  switch (Nc) { case 20480: case 24576: case 36864: case 40960: case 49152: case 61440: case 65536: return true; }
  return false;
}
// cluster-fused form of the same plan (8-CTA clusters, pass A -> pass C through DSMEM, one HBM round trip): verified by CPU
// stepping, NOT YET RUN ON HARDWARE -> only with PFFFT_B200_TILED2D=2
int t2d_cluster_max_active_float(int Nc);
int t2d_cluster_launch_float(int Nc, int sign, const cpx<float>* x, cpx<float>* X, long long batch,
                             const cpx<float>* tables, cudaStream_t st);
// measured on hardware in round 2 (profiles/r02_large_n.md): 16384: 0.48 (4-CTA cluster kernel 0.41, two-launch tiled 0.39) ->
// the default there; 32768: 0.39, 65536: 0.43 -> not faster than the two-launch forms, opt-in (PFFFT_B200_TILED2D=2)
inline bool t2d_cluster_requested(int Nc) { const char* e = getenv("PFFFT_B200_TILED2D"); return e ? atoi(e) == 2 : Nc == 16384; }
inline bool t2d_enabled(int Nc) {
  const char* e = getenv("PFFFT_B200_TILED2D");
  if (e) return atoi(e) != 0;
  return Nc == 16384 || Nc == 32768 || Nc == 65536;
}

// ---- cluster variant (cluster_kernels.cuh, instantiated in cluster.cu): float complex cores (CL*Q) x 4096, rows parked
// in the distributed shared memory of a CL-CTA cluster -> one HBM round trip for 16384 .. 65536 points
// mode 0: strided rows staged by cp.async; 1: rows distributed through DSMEM
bool cluster_shape_exists(int CL, int Q, int mode);
int cluster_max_active_float(int CL, int Q, int mode);          // co-resident clusters on the current device (0: unusable)
int cluster_launch_float(int CL, int Q, int mode, int sign, const cpx<float>* src, cpx<float>* dst, long long batch,
                         const cpx<float>* tw1, const cpx<float>* tw2, const cpx<float>* twP, cudaStream_t st);
// cluster shape for a float plan R x N2 (false = none).  Measured on B200 (profiles/r01b_cluster.md): only 4 x 4096 beats
// the two-pass plan (0.41 vs 0.38 of HBM peak), so it is the one default; the other shapes stay selectable:
//   PFFFT_B200_CLUSTER=0 none, =all every shape that exists (8 x 4096 on 8 CTAs, 16 x 4096 on 8 CTAs x 2 rows);
//   PFFFT_B200_CLUSTER_MODE=0 strided rows staged by cp.async, =1 rows distributed through DSMEM (one row per CTA only);
//   PFFFT_B200_CLUSTER_SHAPE=CLxQ picks the cluster size and rows per CTA for R = CL*Q (4x2, 4x4, 8x2, 16x1 ...);
//   PFFFT_B200_CLUSTER_R16=16 runs 16 x 4096 on 16-CTA clusters; PFFFT_B200_CLUSTER_8192=1 moves 2 x 4096 from the
//   single-CTA kernel to a 2-CTA cluster.
inline bool cluster_choose(int R, int N2, int* CL, int* Q, int* mode) {
  if (N2 != 4096) return false;
  bool all = false;
  if (const char* e = getenv("PFFFT_B200_CLUSTER")) { if (!strcmp(e, "0")) return false; all = !strcmp(e, "all") || !strcmp(e, "1"); }
  int sc = 0;
  if (const char* e = getenv("PFFFT_B200_CLUSTER_MODE")) { sc = atoi(e); if (sc < 0 || sc > 1) sc = 0; }
  int cl = 0, q = 1;
  switch (R) {
    case 2: if (getenv("PFFFT_B200_CLUSTER_8192") && atoi(getenv("PFFFT_B200_CLUSTER_8192"))) cl = 2; break;
    case 4: cl = 4; break;
    case 8: if (all) cl = 8; break;
    case 16:
      if (!all) break;
      if (getenv("PFFFT_B200_CLUSTER_R16") && atoi(getenv("PFFFT_B200_CLUSTER_R16")) == 16) cl = 16;
      else { cl = 8; q = 2; }
      break;
  }
  if (const char* e = getenv("PFFFT_B200_CLUSTER_SHAPE")) {       // "CLxQ", e.g. 4x2: 8 rows on 4-CTA clusters, 2 rows per CTA
    int a = 0, b = 0;
    if (sscanf(e, "%dx%d", &a, &b) == 2 && a * b == R) { cl = a; q = b; }
  }
  if (!cl) return false;
  if (q > 1) sc = 0;
  if (!cluster_shape_exists(cl, q, sc)) return false;
  if (cluster_max_active_float(cl, q, sc) <= 0) return false;
  *CL = cl; *Q = q; *mode = sc;
  return true;
}

// tables of a split plan whose rows run on the CTA core: [tw1: N2][tw2: 16*C][twP: Nc], twP[n1*N2 + k2] = exp(-2 pi i n1 k2 / Nc)
// (row-major copy of the combine twiddles: the single-kernel variants read it with unit stride)
inline size_t split_table_cpx(int Nc, int N2) { return cta_table_cpx(N2) + (size_t)Nc; }
template <typename T> void split_fill_tables(int Nc, int N2, T* dst) {
  cta_fill_tables<T>(N2, dst);
  T* tp = dst + 2 * cta_table_cpx(N2);
  const int R = Nc / N2;
  for (int n1 = 0; n1 < R; ++n1)
    for (int k2 = 0; k2 < N2; ++k2) {
      long double c, sn;
      pfplan::unit_root((long long)n1 * k2, Nc, &c, &sn);
      tp[2 * ((size_t)n1 * N2 + k2)] = (T)c; tp[2 * ((size_t)n1 * N2 + k2) + 1] = (T)sn;
    }
}

// ---- single-kernel variant (k_cta_split): (C, R) pairs that are instantiated; each Nc has exactly one of them
inline bool split_fused_combo(int C, int R) {
  switch (C) {
    case 16: return R == 2;
    case 8: return R == 3 || R == 5 || R == 6;
    case 4: return R == 3 || R == 5 || R == 6 || R == 9 || R == 10 || R == 12;
    case 2: return R == 3 || R == 5 || R == 9 || R == 15;
  }
  return false;
}
template <typename T> inline bool split_fused_ok(int R, int N2);
template <typename T> inline bool split_fused_ok_fwd(int R, int N2) { return split_fused_ok<T>(R, N2); }
template <typename T> inline bool split_fused_ok(int R, int N2) {
  const int C = cta_C_for(N2);
  return C && split_fused_combo(C, R) && (size_t)(R + 1) * N2 * sizeof(cpx<T>) <= 113 * 1024;   // two CTAs per SM
}
template <typename T, int C, int R, int SIGN>
int launch_split_fused_v(Setup<T>* s, const cpx<T>* src, cpx<T>* dst, long long batch, cudaStream_t st) {
  constexpr int MINB = (cta_tpsm<T>() / (16 * C)) < 2 ? 1 : 2;
  auto kern = k_cta_split<T, C, R, SIGN, MINB>;
  const size_t smem = (size_t)(R + 1) * K2<C>::NC * sizeof(cpx<T>);
  static PerDeviceInt cache;                                   // resident CTAs per SM on each device (attribute set first)
  int attr_rc = 0;
  const int per_sm = cache.get(s->device, [&]() -> int {
    if (smem > 48 * 1024) attr_rc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr_rc) return -1;
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 16 * C, smem);
    return n < 1 ? 1 : n;
  });
  if (attr_rc) { set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", (cudaError_t)attr_rc); return attr_rc; }
  long long ctas = batch;
  const long long cap = (long long)s->sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  kern<<<(int)ctas, 16 * C, smem, st>>>(reinterpret_cast<const T*>(src), reinterpret_cast<T*>(dst), batch,
                                         s->tw_fast, s->tw_fast + K2<C>::NC, s->tw_fast + cta_table_cpx(K2<C>::NC));
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T, int SIGN>
int launch_split_fused(Setup<T>* s, const cpx<T>* src, cpx<T>* dst, long long batch, cudaStream_t st) {
  const int C = cta_C_for(s->split_N2), R = s->split_R;
#define PF_SF(c, r) if (C == c && R == r) return launch_split_fused_v<T, c, r, SIGN>(s, src, dst, batch, st);
  PF_SF(16, 2) PF_SF(8, 3) PF_SF(8, 5) PF_SF(8, 6) PF_SF(4, 3) PF_SF(4, 5) PF_SF(4, 6) PF_SF(4, 9) PF_SF(4, 10) PF_SF(4, 12)
  PF_SF(2, 3) PF_SF(2, 5) PF_SF(2, 9) PF_SF(2, 15)
#undef PF_SF
  set_error_msg("split plan marked fused without an instantiated kernel");
  return (int)cudaErrorInvalidValue;
}

// Rows: functor (Setup*, src, rows, batch, stream) -> rc, chosen by the caller for the direction SIGN
template <typename T, int LM, int SM, int SIGN, typename Rows>
int run_split_modes(Setup<T>* s, Rows rows_fn, const XformParams<T>& p, cudaStream_t st) {
  const long long total = p.batch * (long long)s->Nc;
  long long g = (total + 255) / 256; const long long cap = (long long)s->sm_count * 32;
  if (g > cap) g = cap; if (g < 1) g = 1;
  const bool dense_io = p.in_stride == (long long)s->per() && p.out_stride == (long long)s->per();
  // input that already IS the dense complex core (complex canonical, or real time samples read as pairs)
  const bool direct_in = dense_io && p.in_limit < 0 && (LM == L_C_ORD || LM == L_R_TIME) && vec_aligned<T>(p.in);
  const bool direct_out = dense_io && (SM == S_C_ORD || (SM == S_R_TIME && p.out_count >= s->N)) && vec_aligned<T>(p.out);
  const bool one_kernel = !s->d_aux_tables && (s->split_fused || s->split_cluster > 0 || (s->split_t2d && s->split_t2d_cluster));
  const bool need_scratch = !direct_in || !direct_out || !one_kernel;
  std::unique_lock<std::mutex> lock(s->scratch_mu, std::defer_lock);
  if (need_scratch) {
    lock.lock();
    const int rc = scratch_acquire(s, (size_t)p.batch * s->Nc, st);
    if (rc) return rc;
  }
  const cpx<T>* src = reinterpret_cast<const cpx<T>*>(p.in);
  if (!direct_in) {
    k_glob_load<T, LM><<<(int)g, 256, 0, st>>>(p, s->d_scratch[0]);
    count_launch();
    src = s->d_scratch[0];
  }
  // (the fused kernel reads a whole transform before it writes it, so src == dst is fine)
  cpx<T>* dst = direct_out ? reinterpret_cast<cpx<T>*>(p.out) : s->d_scratch[0];
  if (s->d_aux_tables) {                                          // general-radix tiled plan (opt-in)
    int rc = (int)cudaErrorInvalidValue;
    if constexpr (sizeof(T) == 4)
      rc = t2dg_launch_float(s->Nc, SIGN, src, s->d_scratch[1], dst, p.batch, reinterpret_cast<const cpx<float>*>(s->d_aux_tables), s->sm_count, st);
    else
      rc = t2dg_launch_double(s->Nc, SIGN, src, s->d_scratch[1], dst, p.batch, reinterpret_cast<const cpx<double>*>(s->d_aux_tables), s->sm_count, st);
    if (rc) return rc;
  } else if (s->split_t2d && s->split_t2d_cluster) {
    int rc = (int)cudaErrorInvalidValue;
    if constexpr (sizeof(T) == 4)
      rc = t2d_cluster_launch_float(s->Nc, SIGN, src, dst, p.batch, s->tw_fast + split_table_cpx(s->Nc, s->split_N2), st);
    if (rc) return rc;
  } else if (s->split_t2d) {
    int rc = (int)cudaErrorInvalidValue;
    if constexpr (sizeof(T) == 4)
      rc = t2d_launch_float(s->Nc, SIGN, src, s->d_scratch[1], dst, p.batch,
                            s->tw_fast + split_table_cpx(s->Nc, s->split_N2), s->sm_count, st);
    if (rc) return rc;
  } else if (s->split_cluster > 0) {
    int rc = (int)cudaErrorInvalidValue;
    if constexpr (sizeof(T) == 4)
      rc = cluster_launch_float(s->split_cluster, s->split_Q, s->split_mode, SIGN, src, dst, p.batch, s->tw_fast,
                                s->tw_fast + s->split_N2, s->tw_fast + cta_table_cpx(s->split_N2), st);
    if (rc) return rc;
  } else if (s->split_fused) {
    const int rc = launch_split_fused<T, SIGN>(s, src, dst, p.batch, st);
    if (rc) return rc;
  } else {
    { const int rc = rows_fn(s, src, s->d_scratch[1], p.batch, st); if (rc) return rc; }
    { const int rc = split_combine<T, SIGN>(s, s->d_scratch[1], dst, p.batch, st); if (rc) return rc; }
  }
  if (!direct_out) {
    k_glob_store<T, SM><<<(int)g, 256, 0, st>>>(p, s->d_scratch[0]);
    count_launch();
  }
  PF_CUDA_OK(cudaGetLastError());
  if (need_scratch) PF_CUDA_OK(cudaEventRecord(s->scratch_done, st));
  return 0;
}
// RowsF / RowsB: row launchers for the forward / backward direction
template <typename T, typename RowsF, typename RowsB>
int run_split(Setup<T>* s, RowsF rf, RowsB rb, const XformParams<T>& p, int direction, int ordered, cudaStream_t st) {
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? run_split_modes<T, L_C_ORD, S_C_ORD, -1>(s, rf, p, st) : run_split_modes<T, L_C_ORD, S_C_Z, -1>(s, rf, p, st);
    return ordered ? run_split_modes<T, L_C_ORD, S_C_ORD, +1>(s, rb, p, st) : run_split_modes<T, L_C_Z, S_C_ORD, +1>(s, rb, p, st);
  }
  if (fwd) return ordered ? run_split_modes<T, L_R_TIME, S_R_ORD, -1>(s, rf, p, st) : run_split_modes<T, L_R_TIME, S_R_Z, -1>(s, rf, p, st);
  return ordered ? run_split_modes<T, L_R_ORD, S_R_TIME, +1>(s, rb, p, st) : run_split_modes<T, L_R_Z, S_R_TIME, +1>(s, rb, p, st);
}
inline bool is_cta_row_size(int n) { return cta_C_for(n) != 0; }

// hooks for a precision whose only tuned kernels are the CTA ones (double)
template <typename T> struct CtaOnlyHooks {
  // THE decomposition of a core (tables and plan must agree on it): single-kernel form first unless switched off
  static bool decompose(int Nc, int* R, int* N2, bool* fused) {
    if (cta_C_for(Nc)) return false;
    *fused = !getenv("PFFFT_B200_NO_FUSED_SPLIT") && split_choose_fused<T>(Nc, R, N2);
    return *fused || split_choose(Nc, is_cta_row_size, R, N2);
  }
  static int rows_size(int N, int transform) {               // N2 of the split plan, 0 if the size is not split
    int R = 0, N2 = 0; bool fused = false;
    return decompose(transform == XF_REAL ? N / 2 : N, &R, &N2, &fused) ? N2 : 0;
  }
  // double: compile-time-radix CTA kernels for the small and mixed-radix cores (radix_d.cu, round 2b)
  static bool radix_d(int Nc) { if constexpr (sizeof(T) == 8) return radix_core_supported_double(Nc, nullptr); else return false; }
  static size_t extra_table_cpx(int N, int transform) {
    const int Nc = transform == XF_REAL ? N / 2 : N;
    if (ts_wanted_for<T>(Nc) || radix_d(Nc)) return 0;
    const int n2 = rows_size(N, transform);
    return n2 ? split_table_cpx(Nc, n2) : cta_table_cpx(Nc);
  }
  static void fill_extra_table(int N, int transform, T* dst) {
    const int Nc = transform == XF_REAL ? N / 2 : N;
    if (ts_wanted_for<T>(Nc) || radix_d(Nc)) return;
    const int n2 = rows_size(N, transform);
    if (n2) split_fill_tables<T>(Nc, n2, dst); else cta_fill_tables<T>(Nc, dst);
  }
  static bool plan(Setup<T>* s) {
    if (ts_wanted_for<T>(s->Nc)) return ts_plan<T>(s);
    if constexpr (sizeof(T) == 8) {
      const char* nm = "";
      if (radix_core_supported_double(s->Nc, &nm)) { s->fast_variant = 600; s->kernel_name = nm; return true; }
    }
    int R = 0, N2 = 0; bool fused = false;
    if (decompose(s->Nc, &R, &N2, &fused)) {
      if (getenv("PFFFT_B200_NO_SPLIT")) return false;
      s->split_R = R; s->split_N2 = N2; s->split_fused = fused;
      s->fast_variant = 300;
      snprintf(s->name_buf, sizeof(s->name_buf), fused ? "cta_split_%dx%d" : "split_%dx%d", R, N2);
      if constexpr (sizeof(T) == 8) {                            // opt-in tiled plan for doubles (not yet run on hardware)
        int a1 = 0, a2 = 0;
        if (t2dg_wanted(s->Nc, true) && t2dg_shape_for_double(s->Nc, &a1, &a2) && (s->d_aux_tables = t2dg_make_tables_double(s->Nc)) != nullptr) {
          s->split_fused = false;
          snprintf(s->name_buf, sizeof(s->name_buf), "tiled2dg_%dx%d", 16 * a1, 16 * a2);
        }
      }
      s->kernel_name = s->name_buf;
      return true;
    }
    const int C = cta_C_for(s->Nc);
    if (!C || getenv("PFFFT_B200_NO_CTA")) return false;
    s->fast_variant = 100 + C;
    s->kernel_name = cta_name(C);
    return true;
  }
  static int run(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered, cudaStream_t st, const XformOpts& o) {
    if (s->fast_variant == 500) return ts_dispatch<T>(s, in, out, batch, direction, ordered, st, o);
    if constexpr (sizeof(T) == 8) {
      if (s->fast_variant == 600) {                               // compile-time-radix CTA kernels: dense aligned batches
        const bool plain = o.in_stride < 0 && o.out_stride < 0 && o.in_limit < 0 && o.out_count < 0;
        if (!plain || !vec_aligned<T>(in) || !vec_aligned<T>(out)) return -1;
        int lm = 0, sm = 0;
        ts_modes(s->transform, direction, ordered, &lm, &sm);
        return radix_launch_double(s->Nc, lm, sm, direction == DIR_FORWARD ? -1 : +1, in, out, batch, s->tw, s->twr, s->device, s->sm_count, st);
      }
    }
    const XformParams<T> p = make_params(s, in, out, batch, o);
    if (s->fast_variant >= 300)
      return run_split<T>(s, split_rows_cta<T, -1>, split_rows_cta<T, +1>, p, direction, ordered, st);
    return run_cta_any<T>(s, s->fast_variant - 100, p, direction, ordered, st);
  }
};

}  // namespace pf
