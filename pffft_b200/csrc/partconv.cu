// partconv.cu -- uniformly PARTITIONED overlap-save convolution for long filters, under the C-ABI: the
// caller side of pffft_zconvolve_accumulate that the reference describes (include/pffft/pffft.h:182-195: "multiply-accumulate
// of spectra ... useful for partitioned convolution") moved onto the device.
//   taps cut into P partitions of B taps; window W_k = x[kB : kB+2B) (zero padded past the end), S_k = FFT_2B(W_k) (z-domain);
//   Y_k = sum_p S_{k+p} * H_p / (2B)      <- pffft_zconvolve_accumulate applied P times, here ONE launch with the sum over
//                                            p kept in registers (same operation order, bit-identical to P calls)
//   y[kB : (k+1)B) = first B samples of IFFT_2B(Y_k).
// Three launches for any stream length: the forward kernel gathers the overlapping windows itself (batch stride B instead
// of 2B, zero padding by its read limit), the backward kernel stores only the B valid samples of each block straight into y.
// Output convention = pffastconv's (ref src/pffastconv.c:99-106): y[n] = sum_j x[n+j] * taps[F-1-j], n in [0, len-F].
#include <cuda_runtime.h>
#include <vector>
#include "../../include/pffft/pffft_b200.h"
#include "internal_api.h"

struct PFFASTCONVB_Partitioned {
  PFFFT_Setup* st = nullptr;
  int F = 0, B = 0, N = 0, P = 0;
  float* d_H = nullptr;                       // P spectra of N floats, z-domain
  float* d_S = nullptr; size_t S_elems = 0;   // window spectra
  float* d_Y = nullptr; size_t Y_elems = 0;   // block spectra
  float* d_x = nullptr; size_t x_elems = 0;   // host-pointer staging
  float* d_y = nullptr; size_t y_elems = 0;
  cudaStream_t stream = nullptr;
};

namespace pf {
namespace {

int pgrow(float** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) { cudaDeviceSynchronize(); cudaFree(*p); *p = nullptr; *cap = 0; }
  const cudaError_t e = cudaMalloc((void**)p, need * sizeof(float));
  if (e != cudaSuccess) { set_error("partitioned convolution: scratch allocation", e); return (int)e; }
  *cap = need;
  return 0;
}

// Y[k] = sum_{p<P} S[k+p] * H[p] * scaling, real z-domain spectra of `per` floats: one thread per 8-float group
// ([4 re | 4 im]); the reference's operation order per accumulate step (VCPLXMUL, then VMADD into the accumulator,
// src/pffft_priv_impl.h:1534-1630), elements 0 and 4 of group 0 (DC, Nyquist) multiplied as independent reals.
__global__ void __launch_bounds__(256) k_partitioned_mac(const float* __restrict__ S, const float* __restrict__ H, float* __restrict__ Y,
                                                         long long K, int P, int per, float scaling) {
  const int groups = per / 8;
  const long long total = K * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long k = idx / groups;
    const int g = (int)(idx - k * groups);
    float accr[4] = {0.f, 0.f, 0.f, 0.f}, acci[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < P; ++p) {
      const float4 ar = *reinterpret_cast<const float4*>(S + (k + p) * per + 8 * g);
      const float4 ai = *reinterpret_cast<const float4*>(S + (k + p) * per + 8 * g + 4);
      const float4 br = __ldg(reinterpret_cast<const float4*>(H + (long long)p * per + 8 * g));
      const float4 bi = __ldg(reinterpret_cast<const float4*>(H + (long long)p * per + 8 * g + 4));
      const float a_r[4] = {ar.x, ar.y, ar.z, ar.w}, a_i[4] = {ai.x, ai.y, ai.z, ai.w};
      const float b_r[4] = {br.x, br.y, br.z, br.w}, b_i[4] = {bi.x, bi.y, bi.z, bi.w};
      const float c0r = accr[0], c0i = acci[0];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float re = __fadd_rn(__fmul_rn(a_r[l], b_r[l]), -__fmul_rn(a_i[l], b_i[l]));
        const float im = __fadd_rn(__fmul_rn(a_i[l], b_r[l]), __fmul_rn(a_r[l], b_i[l]));
        accr[l] = __fadd_rn(__fmul_rn(re, scaling), accr[l]);
        acci[l] = __fadd_rn(__fmul_rn(im, scaling), acci[l]);
      }
      if (g == 0) {
        accr[0] = __fadd_rn(c0r, __fmul_rn(__fmul_rn(a_r[0], b_r[0]), scaling));
        acci[0] = __fadd_rn(c0i, __fmul_rn(__fmul_rn(a_i[0], b_i[0]), scaling));
      }
    }
    *reinterpret_cast<float4*>(Y + k * per + 8 * g) = make_float4(accr[0], accr[1], accr[2], accr[3]);
    *reinterpret_cast<float4*>(Y + k * per + 8 * g + 4) = make_float4(acci[0], acci[1], acci[2], acci[3]);
  }
}

}  // namespace
}  // namespace pf

using namespace pf;

extern "C" {

PFFFT_EXPORT void pffastconvb_partitioned_destroy(PFFASTCONVB_Partitioned* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (float* p : {c->d_H, c->d_S, c->d_Y, c->d_x, c->d_y}) if (p) cudaFree(p);
  if (c->st) pffft_destroy_setup(c->st);
  delete c;
}

PFFFT_EXPORT PFFASTCONVB_Partitioned* pffastconvb_partitioned_new(const float* taps, int filterLen, int partLen) {
  if (!taps || filterLen <= 0 || partLen < 16 || (partLen & (partLen - 1))) {
    set_error_msg("pffastconvb_partitioned_new: partLen must be a power of two >= 16 (real transforms of 2*partLen points)");
    return nullptr;
  }
  PFFASTCONVB_Partitioned* c = new PFFASTCONVB_Partitioned();
  c->F = filterLen; c->B = partLen; c->N = 2 * partLen; c->P = (filterLen + partLen - 1) / partLen;
  c->st = pffft_new_setup(c->N, PFFFT_REAL);
  if (!c->st) { delete c; return nullptr; }
  // partition p holds hr[pB .. pB+B), hr[j] = taps[F-1-j], time-reversed and placed circularly like the reference's single
  // partition (src/pffastconv.c:99-106)
  const int B = c->B, N = c->N, P = c->P;
  std::vector<float> ht((size_t)P * N, 0.f);
  for (int j = 0; j < filterLen; ++j) {
    const int p = j / B, i = j % B;
    ht[(size_t)p * N + (size_t)((N - i) % N)] = taps[filterLen - 1 - j];
  }
  float* d_t = nullptr;
  bool ok = cudaMalloc((void**)&c->d_H, ht.size() * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void**)&d_t, ht.size() * sizeof(float)) == cudaSuccess &&
            cudaMemcpy(d_t, ht.data(), ht.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
            float_transform_device(c->st, d_t, c->d_H, P, DIR_FORWARD, 0, nullptr, XformOpts()) == 0 &&
            cudaStreamSynchronize(nullptr) == cudaSuccess;
  if (d_t) cudaFree(d_t);
  if (!ok) { set_error("pffastconvb_partitioned_new", cudaGetLastError()); pffastconvb_partitioned_destroy(c); return nullptr; }
  return c;
}

PFFFT_EXPORT int pffastconvb_partitioned_set_stream(PFFASTCONVB_Partitioned* c, void* st) {
  if (!c) return (int)cudaErrorInvalidValue;
  c->stream = (cudaStream_t)st; return 0;
}
PFFFT_EXPORT int pffastconvb_partitioned_partitions(const PFFASTCONVB_Partitioned* c) { return c ? c->P : 0; }

// returns the number of output samples (len - filterLen + 1, or 0), < 0 on error
PFFFT_EXPORT long long pffastconvb_partitioned_apply(PFFASTCONVB_Partitioned* c, const float* input, long long len, float* output) {
  if (!c || !input || !output) { set_error_msg("pffastconvb_partitioned_apply: NULL argument"); return -1; }
  const long long n_out = len - c->F + 1;
  if (n_out <= 0) return 0;
  const bool din = ptr_is_device(input), dout = ptr_is_device(output);
  if (din != dout) { set_error_msg("pffastconvb_partitioned_apply: input and output must both be host or both be device pointers"); return -1; }
  const int B = c->B, N = c->N, P = c->P;
  const long long K = (n_out + B - 1) / B, Kw = K + P - 1;
  cudaStream_t st = c->stream;
  if (pgrow(&c->d_S, &c->S_elems, (size_t)Kw * N) || pgrow(&c->d_Y, &c->Y_elems, (size_t)K * N)) return -1;
  const float* dx = input; float* dy = output;
  if (!din) {
    if (pgrow(&c->d_x, &c->x_elems, (size_t)len + 8) || pgrow(&c->d_y, &c->y_elems, (size_t)n_out + 8)) return -1;
    if (cudaMemcpyAsync(c->d_x, input, (size_t)len * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("partitioned apply: H2D", cudaGetLastError()); return -1; }
    dx = c->d_x; dy = c->d_y;
  }
  // S_k = FFT(x[kB : kB + 2B)), windows gathered by the forward kernel (stride B, zeros past `len`)
  XformOpts fo; fo.in_stride = B; fo.out_stride = N; fo.in_limit = len;
  if (float_transform_device(c->st, dx, c->d_S, Kw, DIR_FORWARD, 0, st, fo)) return -1;
  { long long g = (K * (N / 8) + 255) / 256; const long long cap = 148LL * 16; if (g > cap) g = cap; if (g < 1) g = 1;
    k_partitioned_mac<<<(int)g, 256, 0, st>>>(c->d_S, c->d_H, c->d_Y, K, P, N, 1.0f / (float)N);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) { set_error("partitioned apply: mac launch", cudaGetLastError()); return -1; } }
  // y[kB : kB + B) = first B samples of IFFT(Y_k); the last block may be shorter
  const int last = (int)(n_out - (K - 1) * B);
  XformOpts bo; bo.in_stride = N; bo.out_stride = B; bo.out_count = B;
  if (K > 1 && float_transform_device(c->st, c->d_Y, dy, K - 1, DIR_BACKWARD, 0, st, bo)) return -1;
  bo.out_count = last;
  if (float_transform_device(c->st, c->d_Y + (size_t)(K - 1) * N, dy + (size_t)(K - 1) * B, 1, DIR_BACKWARD, 0, st, bo)) return -1;
  if (!din) {
    if (cudaMemcpyAsync(output, dy, (size_t)n_out * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("partitioned apply: D2H", cudaGetLastError()); return -1; }
  }
  return n_out;
}

}  // extern "C"
