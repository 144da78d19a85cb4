// internal_api.h -- C++ entry points shared between translation units of libpffft_b200.so
// (hidden visibility; not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include "../../include/pffft/pffft_b200.h"
#include "engine.cuh"

namespace pf {
// device-pointer float transform with per-call strides/limits (the engine call behind pffft_transform*)
int float_transform_device(PFFFT_Setup* s, const float* in, float* out, long long batch, int direction, int ordered,
                           cudaStream_t st, const XformOpts& o);
int float_zconvolve_device(PFFFT_Setup* s, const float* a, const float* b, float* ab, float scaling, long long batch,
                           int b_shared, int accumulate, cudaStream_t st);
// tables of a float plan that runs on the 16x16xC CTA kernels (C = 0 when it does not)
struct FloatPlanTables { int C; int sm_count; const cpx<float>* tw1; const cpx<float>* tw2; const cpx<float>* twr; };
FloatPlanTables float_plan_tables(PFFFT_Setup* s);
int float_zreorder_device(PFFFT_Setup* s, const float* in, float* out, long long batch, int direction, cudaStream_t st);
}  // namespace pf
