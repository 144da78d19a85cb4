// fastconv.cu -- pffastconv_* : overlap-save FIR convolution (ref src/pffastconv.c:58-263,
// include/pffft/pffastconv.h:83-180).
//
// Same block algebra and results as the reference's per-block loop
//   memcpy -> pffft_transform(FWD) -> pffft_zconvolve_no_accu(Hf, 1/Nfft) -> pffft_transform(BWD) -> memcpy
// but every block of one pffastconv_apply call is an independent unit of work, so the call becomes
// three batched launches over all blocks at once: the forward kernel reads the overlapping input
// windows straight from the stream (batch stride = outputs per block, zero padding past the end),
// the inverse kernel stores only the valid samples of each block straight into the output.
#include <cuda_runtime.h>
#include <vector>
#include <string.h>
#include <stdlib.h>
#include "../../include/pffft/pffft_b200.h"
#include "internal_api.h"
#include "fastconv_kernels.cuh"

using pf::XformOpts;

struct PFFASTCONV_Setup {
  PFFFT_Setup* st = nullptr;     // real plan of Nfft points
  int filterLen = 0;             // effective taps (2F-1 in single-FFT complex mode, ref pffastconv.c:91-94)
  int Nfft = 0;
  int flags = 0;
  float scale = 0.f;
  int device = 0;
  cudaStream_t stream = nullptr; // device-pointer calls
  float* d_Hf = nullptr;         // z-domain spectrum of the arranged filter
  float* d_Hc = nullptr;         // the same spectrum in canonical order (fused kernel)
  pf::FloatPlanTables tabs{0, 0, nullptr, nullptr, nullptr};
  int fused_ctas_per_sm = 0;
  // scratch, grown on demand (a PFFASTCONV_Setup is single-threaded by contract, pffastconv.h:77-81)
  float* d_spec = nullptr; size_t spec_elems = 0;
  float* d_x = nullptr;    size_t x_elems = 0;     // host input staging / planar split
  float* d_y = nullptr;    size_t y_elems = 0;     // host output staging / planar split
  // host-pointer calls: the stream is cut into pieces that move through three internal streams, so the H2D copy of
  // piece c+1, the kernel of piece c and the D2H copy of piece c-1 overlap (PCIe is full duplex)
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t h2d_done[3] = {nullptr, nullptr, nullptr};
  // stateful stream (pffastconvb_push / _flush): the unconsumed tail of the stream lives on the device between calls
  float* d_carry = nullptr;      // capacity 2*Nfft floats (fewer than Nfft samples ever stay pending)
  long long carry = 0;           // pending (complex) samples
  float* d_stage = nullptr; size_t stage_elems = 0;     // [carry | new input] of one push
  float* d_sout = nullptr;  size_t sout_elems = 0;      // device-side outputs of a host-pointer push
};

namespace {

int grow(float** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) { cudaDeviceSynchronize(); cudaFree(*p); *p = nullptr; *cap = 0; }
  cudaError_t e = cudaMalloc((void**)p, need * sizeof(float));
  if (e != cudaSuccess) { pf::set_error("pffastconv: scratch allocation", e); return (int)e; }
  *cap = need;
  return 0;
}

// interleaved complex stream <-> two planar real streams (CPLX_INP_OUT without SINGLE_FFT runs the
// real convolution on the real and on the imaginary part, ref pffastconv.c:212-247)
__global__ void k_split(const float* __restrict__ x, float* __restrict__ re, float* __restrict__ im, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    re[i] = v.x; im[i] = v.y;
  }
}
__global__ void k_merge(const float* __restrict__ re, const float* __restrict__ im, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float2*>(y)[i] = make_float2(re[i], im[i]);
}
// split/merge for streams whose base is only 4-byte aligned
__global__ void k_split_u(const float* __restrict__ x, float* __restrict__ re, float* __restrict__ im, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    re[i] = x[2 * i]; im[i] = x[2 * i + 1];
  }
}
__global__ void k_merge_u(const float* __restrict__ re, const float* __restrict__ im, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    y[2 * i] = re[i]; y[2 * i + 1] = im[i];
  }
}

using pfplan::BlockPlan;
using pfplan::plan_blocks;

// fused path: one launch for every block of the stream (Nfft = 512*C, C in {2,4,8,16})
// ES = 1: one real stream; ES = 2: interleaved complex stream, both planes in one launch (no split / merge copies)
template <int C, int ES = 1>
int launch_fused(PFFASTCONV_Setup* s, const float* x, long long inputLen, float* y, const BlockPlan& bp, cudaStream_t st) {
  constexpr int MINB = (C == 16) ? 3 : 1024 / (16 * C);
  auto kern = pf::k_fastconv_fused<C, MINB, ES>;
  const size_t smem = 2 * (size_t)pf::K2<C>::NC * sizeof(pf::cf);
  static pf::PerDeviceInt occ;                                   // per device: attribute set, then resident CTAs per SM
  int arc = 0;
  s->fused_ctas_per_sm = occ.get(s->device, [&]() -> int {
    if (smem > 48 * 1024) arc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (arc) return -1;
    int per = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, 16 * C, smem);
    return per < 1 ? 1 : per;
  });
  if (arc) { pf::set_error("pffastconv: cudaFuncSetAttribute", (cudaError_t)arc); return -1; }
  pf::FastconvParams p;
  p.x = x; p.y = y; p.input_len = inputLen; p.n_full = bp.n_full; p.stride = bp.stride;
  p.tail_out = bp.tail_off >= 0 ? bp.tail_out : 0;
  p.scale = s->scale; p.twr = s->tabs.twr; p.Hc = reinterpret_cast<const pf::cf*>(s->d_Hc);
  p.tw1 = s->tabs.tw1; p.tw2 = s->tabs.tw2;
  long long nblk = bp.n_full + (p.tail_out > 0 ? 1 : 0);
  long long ctas = nblk * ES;
  const long long cap = (long long)s->tabs.sm_count * s->fused_ctas_per_sm;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) return 0;
  kern<<<(int)ctas, 16 * C, smem, st>>>(p);
  pf::count_launch();
  if (cudaGetLastError() != cudaSuccess) { pf::set_error("pffastconv: fused launch", cudaGetLastError()); return -1; }
  return 0;
}

// one real stream already on the device: x[0..inputLen) -> y[0..produced)
int conv_stream(PFFASTCONV_Setup* s, const float* x, long long inputLen, float* y, const BlockPlan& bp, cudaStream_t st) {
  const int Nfft = s->Nfft;
  if (s->d_Hc && s->tabs.C) {
    // the tail block follows the full ones at the same stride, so one grid covers both
    switch (s->tabs.C) {
      case 2: return launch_fused<2>(s, x, inputLen, y, bp, st);
      case 4: return launch_fused<4>(s, x, inputLen, y, bp, st);
      case 8: return launch_fused<8>(s, x, inputLen, y, bp, st);
      case 16: return launch_fused<16>(s, x, inputLen, y, bp, st);
    }
  }
  const long long max_blocks = (long long)((size_t)(256u << 20) / ((size_t)Nfft * sizeof(float)));   // <= 256 MiB of spectra in flight
  const long long chunk = max_blocks < 1 ? 1 : max_blocks;
  auto run = [&](long long first_off, long long nblk, int out_count) -> int {
    for (long long b0 = 0; b0 < nblk; b0 += chunk) {
      const long long nb = (nblk - b0 < chunk) ? (nblk - b0) : chunk;
      int rc = grow(&s->d_spec, &s->spec_elems, (size_t)nb * Nfft);
      if (rc) return rc;
      const long long off = first_off + b0 * bp.stride;
      XformOpts fo; fo.in_stride = bp.stride; fo.out_stride = Nfft; fo.in_limit = inputLen - off;
      rc = pf::float_transform_device(s->st, x + off, s->d_spec, nb, pf::DIR_FORWARD, 0, st, fo);
      if (rc) return rc;
      rc = pf::float_zconvolve_device(s->st, s->d_spec, s->d_Hf, s->d_spec, s->scale, nb, 1, 0, st);
      if (rc) return rc;
      XformOpts bo; bo.in_stride = Nfft; bo.out_stride = bp.stride; bo.out_count = out_count;
      rc = pf::float_transform_device(s->st, s->d_spec, y + off, nb, pf::DIR_BACKWARD, 0, st, bo);
      if (rc) return rc;
    }
    return 0;
  };
  int rc = 0;
  if (bp.n_full > 0) rc = run(0, bp.n_full, bp.stride);
  if (!rc && bp.tail_off >= 0) rc = run(bp.tail_off, 1, bp.tail_out);
  return rc;
}

}  // namespace

extern "C" {

PFFASTCONV_EXPORT PFFASTCONV_Setup* pffastconv_new_setup(const float* filterCoeffs, int filterLen, int* blockLen, int flags) {
  if (!filterCoeffs || !blockLen || filterLen <= 0) return nullptr;
  const int cplxFactor = ((flags & PFFASTCONV_CPLX_INP_OUT) && (flags & PFFASTCONV_CPLX_SINGLE_FFT)) ? 2 : 1;
  const int minFftLen = 2 * pffft_simd_size() * pffft_simd_size();
  int Nfft = 2 * pffft_next_power_of_two(filterLen - 1);        // ref pffastconv.c:62
  if (Nfft < minFftLen) Nfft = minFftLen;
  if (flags & PFFASTCONV_CPLX_FILTER) return nullptr;           // ref :71-72
  if (*blockLen > Nfft) Nfft = pffft_next_power_of_two(*blockLen);
  *blockLen = Nfft;                                             // in (complex) samples, ref :80
  Nfft *= cplxFactor;

  PFFASTCONV_Setup* s = new PFFASTCONV_Setup();
  s->st = pffft_new_setup(Nfft, PFFFT_REAL);
  if (!s->st) { delete s; return nullptr; }
  s->device = pffftb_setup_device(s->st);
  s->filterLen = (cplxFactor == 2) ? 2 * filterLen - 1 : filterLen;
  s->Nfft = Nfft; s->flags = flags; s->scale = (float)(1.0 / Nfft);

  // time-reversed taps placed circularly (ref :99-106); zero-stuffed by 2 in single-FFT complex mode
  std::vector<float> ht((size_t)Nfft, 0.f);
  for (int i = 0; i < filterLen; ++i) {
    const float c = (flags & PFFASTCONV_CORRELATION) ? filterCoeffs[i] : filterCoeffs[filterLen - 1 - i];
    ht[(size_t)((Nfft - cplxFactor * i) & (Nfft - 1))] = c;
  }
  bool ok = cudaMalloc((void**)&s->d_Hf, (size_t)Nfft * sizeof(float)) == cudaSuccess;
  float* d_tmp = nullptr;
  ok = ok && cudaMalloc((void**)&d_tmp, (size_t)Nfft * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMemcpy(d_tmp, ht.data(), (size_t)Nfft * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    ok = pf::float_transform_device(s->st, d_tmp, s->d_Hf, 1, pf::DIR_FORWARD, 0, nullptr, XformOpts()) == 0;   // ref :108
    ok = ok && cudaStreamSynchronize(nullptr) == cudaSuccess;
  }
  // canonical copy of the spectrum for the fused single-kernel path
  s->tabs = pf::float_plan_tables(s->st);
  if (ok && s->tabs.C && getenv("PFFFT_B200_NO_FUSED_CONV") == nullptr) {
    ok = cudaMalloc((void**)&s->d_Hc, (size_t)Nfft * sizeof(float)) == cudaSuccess &&
         pf::float_zreorder_device(s->st, s->d_Hf, s->d_Hc, 1, pf::DIR_FORWARD, nullptr) == 0 &&
         cudaStreamSynchronize(nullptr) == cudaSuccess;
  }
  if (d_tmp) cudaFree(d_tmp);
  if (!ok) { pf::set_error("pffastconv_new_setup", cudaGetLastError()); pffastconv_destroy_setup(s); return nullptr; }
  return s;
}

PFFASTCONV_EXPORT void pffastconv_destroy_setup(PFFASTCONV_Setup* s) {
  if (!s) return;
  cudaDeviceSynchronize();
  if (s->d_Hf) cudaFree(s->d_Hf);
  if (s->d_Hc) cudaFree(s->d_Hc);
  if (s->d_spec) cudaFree(s->d_spec);
  if (s->d_x) cudaFree(s->d_x);
  if (s->d_y) cudaFree(s->d_y);
  if (s->d_carry) cudaFree(s->d_carry);
  if (s->d_stage) cudaFree(s->d_stage);
  if (s->d_sout) cudaFree(s->d_sout);
  for (int k = 0; k < 3; ++k) { if (s->hs[k]) cudaStreamDestroy(s->hs[k]); if (s->h2d_done[k]) cudaEventDestroy(s->h2d_done[k]); }
  pffft_destroy_setup(s->st);
  delete s;
}

PFFASTCONV_EXPORT int pffastconv_apply(PFFASTCONV_Setup* s, const float* input, int cplxInputLen, float* output, int applyFlush) {
  if (!s || !input || !output || cplxInputLen <= 0) return 0;
  const int flags = s->flags;
  const bool cplx = (flags & PFFASTCONV_CPLX_INP_OUT) != 0;
  const int cplxFactor = (cplx && (flags & PFFASTCONV_CPLX_SINGLE_FFT)) ? 2 : 1;
  const long long inputLen = (long long)cplxFactor * cplxInputLen;           // length of the real stream(s)
  const BlockPlan bp = plan_blocks(inputLen, s->Nfft, s->filterLen, applyFlush, cplxFactor == 2);
  if (bp.produced <= 0) return 0;

  const bool din = pf::ptr_is_device(input), dout = pf::ptr_is_device(output);
  if (din != dout) { pf::set_error_msg("pffastconv_apply: input and output must both be host or both be device pointers"); return 0; }
  cudaStream_t st = s->stream;
  const size_t in_floats = (size_t)(cplx ? 2 : 1) * (size_t)cplxInputLen;   // floats in `input`
  const size_t out_floats = (size_t)(cplx && cplxFactor == 1 ? 2 : 1) * (size_t)bp.produced;   // floats written to `output`

  const float* dx = input;
  float* dy = output;
  const bool two_planes = cplx && cplxFactor == 1;
  int rc = 0;
  // staging: host pointers need a device copy; the two-FFT complex mode needs planar streams
  size_t need_x = 0, need_y = 0;
  if (!din) { need_x += in_floats; need_y += out_floats; }
  if (two_planes) { need_x += 2 * (size_t)cplxInputLen; need_y += 2 * (size_t)bp.produced; }
  if ((rc = grow(&s->d_x, &s->x_elems, need_x + 8))) return 0;
  if ((rc = grow(&s->d_y, &s->y_elems, need_y + 8))) return 0;
  float* x_planes = s->d_x;
  float* y_planes = s->d_y;
  if (!din && !two_planes && s->d_Hc && s->tabs.C && getenv("PFFFT_B200_CONV_NO_PIPELINE") == nullptr) {
    // ---- pipelined host path (fused kernel only: it needs no per-call scratch, so pieces may overlap freely)
    float* hx = s->d_x; float* hy = s->d_y;
    hx += (4 - ((uintptr_t)hx / sizeof(float)) % 4) % 4;
    hy += (4 - ((uintptr_t)hy / sizeof(float)) % 4) % 4;
    for (int k = 0; k < 3; ++k) {
      if (!s->hs[k] && cudaStreamCreateWithFlags(&s->hs[k], cudaStreamNonBlocking) != cudaSuccess) { pf::set_error("pffastconv_apply: stream", cudaGetLastError()); return 0; }
      if (!s->h2d_done[k] && cudaEventCreateWithFlags(&s->h2d_done[k], cudaEventDisableTiming) != cudaSuccess) { pf::set_error("pffastconv_apply: event", cudaGetLastError()); return 0; }
    }
    const long long nblk = bp.n_full + (bp.tail_off >= 0 ? 1 : 0);
    size_t piece_bytes = (size_t)8 << 20;                      // ~8 MiB of new input per piece (PFFFT_B200_CONV_PIECE_KB overrides)
    if (const char* e = getenv("PFFFT_B200_CONV_PIECE_KB")) { const long kb = atol(e); if (kb > 0) piece_bytes = (size_t)kb << 10; }
    long long per_piece = (long long)(piece_bytes / ((size_t)bp.stride * sizeof(float)));
    if (per_piece < 1) per_piece = 1;
    long long copied = 0;                                      // input floats already on their way to the device
    int k = 0;
    bool okp = true;
    for (long long b0 = 0; b0 < nblk && okp; b0 += per_piece, k = (k + 1) % 3) {
      const long long b1 = (b0 + per_piece < nblk) ? b0 + per_piece : nblk;
      const long long off0 = b0 * bp.stride;
      long long in_end = (b1 - 1) * bp.stride + s->Nfft;       // last sample the piece reads (window of its last block)
      if (in_end > inputLen) in_end = inputLen;
      cudaStream_t ps = s->hs[k];
      if (in_end > copied) {
        okp = cudaMemcpyAsync(hx + copied, input + copied, (size_t)(in_end - copied) * sizeof(float), cudaMemcpyHostToDevice, ps) == cudaSuccess;
        copied = in_end;
      }
      okp = okp && cudaEventRecord(s->h2d_done[k], ps) == cudaSuccess;
      // the head of this piece's first window was copied by the previous piece, on another stream
      if (b0 > 0) {
        okp = okp && cudaStreamWaitEvent(ps, s->h2d_done[(k + 2) % 3], 0) == cudaSuccess;
        okp = okp && cudaStreamWaitEvent(ps, s->h2d_done[(k + 1) % 3], 0) == cudaSuccess;   // (one-block pieces of huge Nfft)
      }
      BlockPlan sub;
      sub.stride = bp.stride;
      const bool has_tail = bp.tail_off >= 0 && b1 == nblk;
      sub.n_full = (b1 - b0) - (has_tail ? 1 : 0);
      sub.tail_off = has_tail ? bp.tail_off - off0 : -1;
      sub.tail_out = has_tail ? bp.tail_out : 0;
      const long long out_count = sub.n_full * bp.stride + sub.tail_out;
      okp = okp && conv_stream(s, hx + off0, inputLen - off0, hy + off0, sub, ps) == 0;
      okp = okp && cudaMemcpyAsync(output + off0, hy + off0, (size_t)out_count * sizeof(float), cudaMemcpyDeviceToHost, ps) == cudaSuccess;
    }
    for (int q = 0; q < 3; ++q) okp = (cudaStreamSynchronize(s->hs[q]) == cudaSuccess) && okp;
    if (!okp) { pf::set_error("pffastconv_apply: pipelined host path", cudaGetLastError()); return 0; }
    return (int)(bp.produced / cplxFactor);
  }
  if (!din) {
    float* hx = s->d_x + (two_planes ? 2 * (size_t)cplxInputLen : 0);
    float* hy = s->d_y + (two_planes ? 2 * (size_t)bp.produced : 0);
    hx += (4 - ((uintptr_t)hx / sizeof(float)) % 4) % 4;      // keep 16-byte alignment of the staged streams
    hy += (4 - ((uintptr_t)hy / sizeof(float)) % 4) % 4;
    if (cudaMemcpyAsync(hx, input, in_floats * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) { pf::set_error("pffastconv_apply: H2D", cudaGetLastError()); return 0; }
    dx = hx; dy = hy;
  }

  if (!two_planes) {
    rc = conv_stream(s, dx, inputLen, dy, bp, st);
  } else if (s->d_Hc && s->tabs.C && getenv("PFFFT_B200_CONV_SPLIT_PLANES") == nullptr) {
    // two-FFT complex mode on the fused kernel: every (block, plane) pair is a work item of ONE launch; the planes are
    // read and written in place in the interleaved streams
    switch (s->tabs.C) {
      case 2: rc = launch_fused<2, 2>(s, dx, inputLen, dy, bp, st); break;
      case 4: rc = launch_fused<4, 2>(s, dx, inputLen, dy, bp, st); break;
      case 8: rc = launch_fused<8, 2>(s, dx, inputLen, dy, bp, st); break;
      default: rc = launch_fused<16, 2>(s, dx, inputLen, dy, bp, st); break;
    }
  } else {
    const long long n = cplxInputLen;
    const long long gcap = (long long)s->tabs.sm_count * 16;
    const int thr = 256; long long g = (n + thr - 1) / thr; if (g > gcap) g = gcap;
    float* xr = x_planes; float* xi = x_planes + n;
    float* yr = y_planes; float* yi = y_planes + bp.produced;
    if (((uintptr_t)dx & 7) == 0) k_split<<<(int)g, thr, 0, st>>>(dx, xr, xi, n); else k_split_u<<<(int)g, thr, 0, st>>>(dx, xr, xi, n);
    pf::count_launch();
    rc = conv_stream(s, xr, inputLen, yr, bp, st);
    if (!rc) rc = conv_stream(s, xi, inputLen, yi, bp, st);
    if (!rc) {
      const long long m = bp.produced; long long g2 = (m + thr - 1) / thr; if (g2 > gcap) g2 = gcap;
      if (((uintptr_t)dy & 7) == 0) k_merge<<<(int)g2, thr, 0, st>>>(yr, yi, dy, m); else k_merge_u<<<(int)g2, thr, 0, st>>>(yr, yi, dy, m);
      pf::count_launch();
    }
  }
  if (rc) return 0;
  if (!din) {
    if (cudaMemcpyAsync(output, dy, out_floats * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { pf::set_error("pffastconv_apply: D2H", cudaGetLastError()); return 0; }
  }
  if (cudaGetLastError() != cudaSuccess) return 0;
  return (int)(bp.produced / cplxFactor);                      // ref :201, :262
}

// ---- stateful stream: the caller of pffastconv_apply's "re-feed what was not consumed" contract
// (include/pffft/pffastconv.h:160-171; src/pffastconv.c:201, :262) moved under the boundary.  Blocks of successive pushes
// start at multiples of the block stride counted from the beginning of the stream -- exactly where one call over the whole
// stream puts them -- so any chunking followed by one flush yields bit-identical samples.
static int stream_step(PFFASTCONV_Setup* s, const float* input, long long n, float* output, long long out_capacity, int flush) {
  const int w = (s->flags & PFFASTCONV_CPLX_INP_OUT) ? 2 : 1;              // floats per (complex) sample
  const bool dout = output && pf::ptr_is_device(output);
  if (n > 0 && pf::ptr_is_device(input) != dout) { pf::set_error_msg("pffastconvb_push: input and output must both be host or both be device pointers"); return -1; }
  const long long total = s->carry + n;
  if (total > 0x7fffffffLL) { pf::set_error_msg("pffastconvb_push: more than 2^31 pending samples"); return -1; }
  cudaStream_t st = s->stream;
  if (!s->d_carry && cudaMalloc((void**)&s->d_carry, (size_t)2 * s->Nfft * sizeof(float) + 64) != cudaSuccess) { pf::set_error("pffastconvb_push: carry buffer", cudaGetLastError()); return -1; }
  if (grow(&s->d_stage, &s->stage_elems, (size_t)total * w + 8)) return -1;
  if (s->carry && cudaMemcpyAsync(s->d_stage, s->d_carry, (size_t)s->carry * w * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
  if (n > 0 && cudaMemcpyAsync(s->d_stage + s->carry * w, input, (size_t)n * w * sizeof(float), cudaMemcpyDefault, st) != cudaSuccess) { pf::set_error("pffastconvb_push: staging copy", cudaGetLastError()); return -1; }
  // how many samples this step will produce (same algebra as pffastconv_apply)
  const int cplxFactor = (w == 2 && (s->flags & PFFASTCONV_CPLX_SINGLE_FFT)) ? 2 : 1;
  const BlockPlan bp = plan_blocks((long long)cplxFactor * total, s->Nfft, s->filterLen, flush, cplxFactor == 2);
  const long long produced = bp.produced / cplxFactor;
  if (produced > out_capacity) { pf::set_error_msg("pffastconvb_push: output capacity too small for the samples this call produces"); return -1; }
  if (produced > 0) {
    float* dy = output;
    if (!dout) { if (grow(&s->d_sout, &s->sout_elems, (size_t)produced * w + 8)) return -1; dy = s->d_sout; }
    const int got = pffastconv_apply(s, s->d_stage, (int)total, dy, flush);
    if (got != (int)produced) { if (!*pffftb_last_error()) pf::set_error_msg("pffastconvb_push: block algebra mismatch"); return -1; }
    if (!dout && cudaMemcpyAsync(output, dy, (size_t)produced * w * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) { pf::set_error("pffastconvb_push: D2H", cudaGetLastError()); return -1; }
  }
  const long long rest = total - produced;                                 // < Nfft after a step without flush; F-1 after a flush
  if (rest * w > 2LL * s->Nfft) { pf::set_error_msg("pffastconvb_push: internal carry overflow"); return -1; }
  if (rest > 0 && cudaMemcpyAsync(s->d_carry, s->d_stage + produced * w, (size_t)rest * w * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
  s->carry = rest;
  if (!dout && cudaStreamSynchronize(st) != cudaSuccess) { pf::set_error("pffastconvb_push: synchronize", cudaGetLastError()); return -1; }
  return (int)produced;
}
PFFFT_EXPORT int pffastconvb_push(PFFASTCONV_Setup* s, const float* input, int cplxInputLen, float* output, int outputCapacity) {
  if (!s || cplxInputLen < 0 || (cplxInputLen > 0 && !input)) { pf::set_error_msg("pffastconvb_push: bad argument"); return -1; }
  return stream_step(s, input, cplxInputLen, output, outputCapacity, 0);
}
PFFFT_EXPORT int pffastconvb_flush(PFFASTCONV_Setup* s, float* output, int outputCapacity) {
  if (!s) { pf::set_error_msg("pffastconvb_flush: bad argument"); return -1; }
  return stream_step(s, nullptr, 0, output, outputCapacity, 1);
}
PFFFT_EXPORT int pffastconvb_pending(const PFFASTCONV_Setup* s) { return s ? (int)s->carry : 0; }
PFFFT_EXPORT void pffastconvb_reset(PFFASTCONV_Setup* s) { if (s) s->carry = 0; }

PFFASTCONV_EXPORT void* pffastconv_malloc(size_t nb) { return pffft_aligned_malloc(nb); }
PFFASTCONV_EXPORT void pffastconv_free(void* p) { pffft_aligned_free(p); }
PFFASTCONV_EXPORT int pffastconv_simd_size(void) { return pffft_simd_size(); }
PFFFT_EXPORT int pffastconvb_set_stream(PFFASTCONV_Setup* s, void* st) {
  if (!s) return (int)cudaErrorInvalidValue;
  s->stream = (cudaStream_t)st; return 0;
}

}  // extern "C"
