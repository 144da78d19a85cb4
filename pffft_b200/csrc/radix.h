// radix.h -- host interface of the compile-time-radix CTA kernels (radix_kernels.cuh; instantiated in radix_a.cu / radix_b.cu)
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace pf {
// true when the float complex core Nc has a compile-time-radix plan; name = "radix_20x20x10" ...
bool radix_core_supported(int Nc, const char** name);
// dense batch (2*Nc floats per transform, 8-byte aligned); lm/sm = LoadMode/StoreMode of generic_kernels.cuh;
// tw = exp(-2 pi i k/Nc), twr = exp(-2 pi i k/(2 Nc)) for real plans.  Returns -1 when (Nc, lm, sm) is not instantiated.
int radix_launch_float(int Nc, int lm, int sm, int sign, const float* in, float* out, long long batch, const cpx<float>* tw,
                       const cpx<float>* twr, int device, int sm_count, cudaStream_t st);
// double precision (radix_d.cu): the cores below 512 that the double CTA kernels (512 ... 4096) do not cover
bool radix_core_supported_double(int Nc, const char** name);
int radix_launch_double(int Nc, int lm, int sm, int sign, const double* in, double* out, long long batch, const cpx<double>* tw,
                        const cpx<double>* twr, int device, int sm_count, cudaStream_t st);
}  // namespace pf
