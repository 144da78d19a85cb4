// fast.h -- host interface of the warp-per-transform float kernels (fast_kernels.cuh), instantiated in two translation units:
// fast_pk.cu is compiled with the packed f32x2 arithmetic of common.cuh, fast_sc.cu with scalar arithmetic (PF_NO_PACKED_F32).
// Packed FADD2/FMUL2/FFMA2 issue at half the rate of the scalar forms (tools/ubench_fp32.cu: same flops per clock, half the
// issue slots), so they pay where a kernel is bound by instruction issue and cost a little where it is bound by latency:
// measured per kernel (profiles/r02b_packed.md), `fast_packed` is the resulting assignment.  Internal to libpffft_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace pf {
struct FastCtx { int device, sm_count; const cpx<float>* tw_fast; const cpx<float>* twr; };
enum FastKind { FK_C1024 = 0, FK_WSMALL = 1, FK_WMIXED = 2 };

// which build serves (kind, R2, z-domain input/output, real): true = packed arithmetic
constexpr bool fast_packed(int kind, int R2, bool zin, bool zout, bool real) {
#ifdef PF_FAST_FORCE                                          // A/B builds: everything on one side
  return PF_FAST_FORCE != 0;
#endif
  if (kind == FK_C1024) return zin || zout;                 // ordered: 0.98 scalar / 0.97 packed; z-domain: 0.90 / 0.96
  if (kind == FK_WSMALL) return R2 == 4;                    // 128c: 0.87 / 0.97; 32c, 64c, 256c: equal (256c z-domain: 0.96 / 0.94)
  if (real) return true;                                    // real 64 ... 1920: +3 ... +17 % packed (backward 512: equal)
  if (zin || zout) return R2 >= 10;
  return R2 >= 10;                                          // 96c: 0.95 / 0.84, 288c: 0.92 / 0.85, 320c: 0.86 / 0.88, 480c: 0.84 / 0.91,
                                                            // 768c: 0.85 / 0.96, 864c: 0.81 / 0.94, 960c: 0.72 / 0.91
}

// c1024: `variant` = V_* of api_float.cu; every call returns -1 when the combination is not instantiated
int fast_c1024_pk(const FastCtx&, int variant, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st);
int fast_c1024_sc(const FastCtx&, int variant, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st);
int fast_wsmall_pk(const FastCtx&, int R2, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st);
int fast_wsmall_sc(const FastCtx&, int R2, int sign, bool zin, bool zout, const float* in, float* out, long long batch, cudaStream_t st);
int fast_wmixed_pk(const FastCtx&, int R2, int sign, bool zin, bool zout, bool real, const float* in, float* out, long long batch, cudaStream_t st, int grp);
int fast_wmixed_sc(const FastCtx&, int R2, int sign, bool zin, bool zout, bool real, const float* in, float* out, long long batch, cudaStream_t st, int grp);
}  // namespace pf
