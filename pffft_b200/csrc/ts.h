// ts.h -- host interface of the tiled Stockham pipeline (ts.cu / ts_kernels.cuh): large complex cores as 2-4 passes of radix
// 16*A whose intermediates stay in L2.  Internal to libpffft_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace pf {
struct TsPlanHost;
// Nc = prod_{i<P} 16*A[i] (2 <= P <= 4, A from the register DFT library), false when Nc has no such factorisation
bool ts_factorize(int Nc, int* P, int* A);
// plan resources on the current device (radix tables, ring buffers, counters); nullptr when Nc is not factorisable or an
// allocation fails (pffftb_last_error says which)
TsPlanHost* ts_create(int N, int Nc, bool dbl, int device, int sm_count);
void ts_destroy(TsPlanHost*);
const char* ts_name(const TsPlanHost*);
// dense batch, `in`/`out` aligned to one complex word; lm/sm are LoadMode/StoreMode values (generic_kernels.cuh);
// tw = exp(-2 pi i k/Nc) (Nc entries), twr = exp(-2 pi i k/N) (real plans)
template <typename T>
int ts_run(TsPlanHost*, const T* in, T* out, long long batch, int sign, int lm, int sm, const cpx<T>* tw, const cpx<T>* twr,
           cudaStream_t st);
}  // namespace pf
