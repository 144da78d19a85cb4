// ts.cu -- host side of the tiled Stockham pipeline (ts_kernels.cuh): factorisation of the core length into radices 16*A,
// per-plan device resources (radix tables, L2-resident ring buffers, dependency counters) and the launcher.
// Own translation unit so the C-ABI units stay small; float and double.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "internal_api.h"
#include "ts.h"
#include "ts_kernels.cuh"
#include "ts_plan.h"

namespace pf {

struct TsPlanHost {
  int Nc = 0, N = 0, device = 0, sm_count = 0;
  bool dbl = false;
  int P = 0, A[4] = {0, 0, 0, 0}, tw_off[4] = {0, 0, 0, 0}, twR_entries = 0;
  void* d_twR = nullptr;
  void* d_ring[kTsMaxRings] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int nrings = 0;
  unsigned* d_counters = nullptr;
  size_t counter_bytes = 0;
  int grid = 0, lag = 0, ring_slots = 1;
  size_t smem = 0;
  bool wmode = false;                  // warp-sized work items (tsw_kernels.cuh)
  int twL_entries = 0;                 // wmode: entries of the last pass's exponent table (appended to d_twR)
  bool pre = false;                    // input prefetch variant available (float)
  int grid_v[2] = {0, 0};              // co-resident grid and shared memory of the variant without / with prefetch
  size_t smem_v[2] = {0, 0};
  std::mutex mu;                       // rings and counters serve one launch at a time
  cudaEvent_t done = nullptr;
  char name[48] = {0};
};

// Kernel variants.  float: FOUR resident CTAs per SM (64 registers since the packed-arithmetic build; measured 0.42 / 0.39 /
// 0.38 of the HBM roofline at 16384 / 32768 / 65536 against 0.38 / 0.37 / 0.35 with three and 0.29 / 0.29 / 0.28 with two:
// the pipeline is bound by latency, i.e. by resident warps).  PFFFT_B200_TS_MINB=2|3 selects the other builds,
// PFFFT_B200_TS_PRE=1 (with MINB 2|3) the cp.async input prefetch -- measured SLOWER (16384: 0.34 against 0.38: the
// second 32 KB buffer leaves 12 KB of L1 for the twiddle tables).  Tuning knobs, read once.  double: 2 CTAs, no prefetch.
template <typename T> struct TsKernels {
  using Kern = void (*)(const TsParams<T>);
  static int minb() {
    static const int v = [] { if (sizeof(T) == 8) return 2; const char* e = getenv("PFFFT_B200_TS_MINB"); const int m = e ? atoi(e) : 4; return (m == 2 || m == 3) ? m : 4; }();
    return v;
  }
  static bool pre() {
    static const bool v = [] { if (sizeof(T) == 8 || minb() == 4) return false; const char* e = getenv("PFFFT_B200_TS_PRE"); return e && atoi(e) == 1; }();
    return v;
  }
  // exchange tile [+ staging buffer] + per-radix tables (entries rounded up to 128 bytes)
  static size_t smem(int twR_entries, bool with_pre) {
    return ((size_t)16 * 256 * (with_pre ? 2 : 1) + (size_t)((twR_entries + 15) / 16) * 16) * sizeof(cpx<T>);
  }
  template <int SIGN> static Kern kern(bool with_pre) {
    if constexpr (sizeof(T) == 8) return (Kern)k_ts_pipeline<T, SIGN, 2, false>;
    else {
      if (minb() == 4) return (Kern)k_ts_pipeline<T, SIGN, 4, false>;
      if (minb() == 2) return with_pre ? (Kern)k_ts_pipeline<T, SIGN, 2, true> : (Kern)k_ts_pipeline<T, SIGN, 2, false>;
      return with_pre ? (Kern)k_ts_pipeline<T, SIGN, 3, true> : (Kern)k_ts_pipeline<T, SIGN, 3, false>;
    }
  }
};

template <typename T> static int ts_prepare_kernels(TsPlanHost* h) {
  static PerDeviceInt attr[4];
  int k = 0;
  for (int with_pre = 0; with_pre <= (TsKernels<T>::pre() ? 1 : 0); ++with_pre) {
    const size_t smem = TsKernels<T>::smem(h->twR_entries, with_pre != 0);
    const size_t smem_max = TsKernels<T>::smem(4 * 256, with_pre != 0);     // the attribute is set once per device: largest plan
    { const int rc = ensure_dyn_smem(attr[k++], h->device, TsKernels<T>::template kern<-1>(with_pre != 0), smem_max); if (rc) return rc; }
    { const int rc = ensure_dyn_smem(attr[k++], h->device, TsKernels<T>::template kern<+1>(with_pre != 0), smem_max); if (rc) return rc; }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, TsKernels<T>::template kern<-1>(with_pre != 0), kTsThreads, smem) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n < 1) n = 1;
    h->grid_v[with_pre] = n * h->sm_count;
    h->smem_v[with_pre] = smem;
  }
  h->pre = TsKernels<T>::pre();
  h->grid = h->grid_v[h->pre ? 1 : 0];
  h->smem = h->smem_v[h->pre ? 1 : 0];
  return 0;
}

// warp-sized work items: float only; 4 CTAs of 4 warps per SM (128 registers per thread), PFFFT_B200_TSW_MINB=2..6 overrides
struct TswKernels {
  using Kern = void (*)(const TsParams<float>, int);
  static int minb() {
    static const int v = [] { const char* e = getenv("PFFFT_B200_TSW_MINB"); const int m = e ? atoi(e) : 4; return (m >= 2 && m <= 6) ? m : 4; }();
    return v;
  }
  static size_t smem(int twR_entries, int twL_entries) {
    return ((size_t)((twR_entries + 15) & ~15) + (size_t)((twL_entries + 15) & ~15) + (size_t)kTswWarps * kTswTileMax) * sizeof(cpx<float>);
  }
  template <int SIGN> static Kern kern() {
    switch (minb()) {
      case 2: return (Kern)k_tsw_pipeline<float, SIGN, 2>;
      case 3: return (Kern)k_tsw_pipeline<float, SIGN, 3>;
      case 5: return (Kern)k_tsw_pipeline<float, SIGN, 5>;
      case 6: return (Kern)k_tsw_pipeline<float, SIGN, 6>;
      default: return (Kern)k_tsw_pipeline<float, SIGN, 4>;
    }
  }
};
static int tsw_prepare_kernels(TsPlanHost* h) {
  static PerDeviceInt attr_f, attr_b;
  const size_t smem = TswKernels::smem(h->twR_entries, h->twL_entries);
  const size_t smem_max = TswKernels::smem(4 * 256, 256);
  { const int rc = ensure_dyn_smem(attr_f, h->device, TswKernels::kern<-1>(), smem_max); if (rc) return rc; }
  { const int rc = ensure_dyn_smem(attr_b, h->device, TswKernels::kern<+1>(), smem_max); if (rc) return rc; }
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, TswKernels::kern<-1>(), kTswWarps * 32, smem) != cudaSuccess) { cudaGetLastError(); n = 0; }
  if (n < 1) n = 1;
  h->grid = n * h->sm_count;
  h->smem = smem;
  return 0;
}

template <typename T> static bool ts_fill_radix_tables(TsPlanHost* h) {
  std::vector<T> host = ts_radix_tables<T>(h->P, h->A, h->tw_off);
  h->twR_entries = (int)(host.size() / 2);
  if (h->wmode) {                                                 // + exponent table of the last pass
    const std::vector<T> last = tsw_last_table<T>(ts_radix(h->A[h->P - 1]));
    h->twL_entries = (int)(last.size() / 2);
    host.insert(host.end(), last.begin(), last.end());
  }
  if (cudaMalloc(&h->d_twR, host.size() * sizeof(T)) != cudaSuccess) return false;
  return cudaMemcpy(h->d_twR, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

void ts_destroy(TsPlanHost* h) {
  if (!h) return;
  if (h->done) { cudaEventSynchronize(h->done); cudaEventDestroy(h->done); }
  if (h->d_twR) cudaFree(h->d_twR);
  for (int i = 0; i < kTsMaxRings; ++i) if (h->d_ring[i]) cudaFree(h->d_ring[i]);
  if (h->d_counters) cudaFree(h->d_counters);
  delete h;
}

TsPlanHost* ts_create(int N, int Nc, bool dbl, int device, int sm_count) {
  int P = 0, A[4];
  if (!ts_factorize(Nc, &P, A)) return nullptr;
  TsPlanHost* h = new TsPlanHost();
  h->N = N; h->Nc = Nc; h->dbl = dbl; h->device = device; h->sm_count = sm_count; h->P = P;
  memcpy(h->A, A, sizeof(int) * P);
  const size_t csz = dbl ? sizeof(cpx<double>) : sizeof(cpx<float>);
  int amax = 0; long long group = 0;
  for (int i = 0; i < P; ++i) { if (A[i] > amax) amax = A[i]; group += ts_tiles(Nc, A[i]); }
  (void)amax;
  // warp-sized work items (tsw_kernels.cuh) for power-of-two float plans: opt-in, PFFFT_B200_TSW=1 -- correct on hardware,
  // but measured at 0.30-0.32 of the roofline at 16384 ... 65536 against 0.41-0.45 for the CTA-sized items
  { const char* e = getenv("PFFFT_B200_TSW"); h->wmode = !dbl && tsw_plan_ok(P, A) && e && atoi(e) == 1; }
  if (h->wmode) { group = 0; for (int i = 0; i < P; ++i) group += ts_tiles(Nc, A[i], true); }
  bool ok = dbl ? ts_fill_radix_tables<double>(h) : ts_fill_radix_tables<float>(h);
  ok = ok && (h->wmode ? tsw_prepare_kernels(h) : (dbl ? ts_prepare_kernels<double>(h) : ts_prepare_kernels<float>(h))) == 0;
  // pipeline depth.  CTA kernel (interleaved order): pass i+1 of a transform is handed out `lag` groups after pass i -- about
  // 1.5 grid-fulls of work items later, so its input is complete (no spinning) and still in L2; rings hold 2*lag+1
  // transforms so a slot's previous occupant has long been consumed when it is overwritten.  Warp kernel (stage-specialised
  // workers): a pass holds workers / (P * tiles) transforms in flight and its ring needs about twice that.  Either way the
  // rings must stay L2 resident to pay: at most ~48 MB over the rings a complex ordered call touches.
  h->nrings = P;                                                  // P-1 between the passes + one for a pre-/post-stage
  const size_t tb = (size_t)Nc * csz;
  const size_t budget = (size_t)(getenv("PFFFT_B200_TS_RING_MB") ? atoi(getenv("PFFFT_B200_TS_RING_MB")) : 48) << 20;
  const size_t hot = (size_t)(P > 2 ? P - 1 : 1);                 // (the extra ring serves pre/post stages)
  long long lag;
  if (h->wmode) {
    const long long tiles0 = ts_tiles(Nc, A[0], true);
    lag = ((long long)h->grid * kTswWarps / P + tiles0 - 1) / tiles0 + 1;
  } else {
    lag = (3LL * h->grid / 2 + group - 1) / group;
    if (lag < 1) lag = 1;
  }
  const long long fit = ((long long)(budget / (hot * tb)) - 1) / 2;
  if (lag > fit) lag = fit >= 1 ? fit : (((size_t)h->nrings * 3 * tb <= ((size_t)512 << 20)) ? 1 : 0);
  if (const char* e = getenv("PFFFT_B200_TS_LAG")) { const long long v = atoll(e); if (v >= 0 && v < 4096) lag = v; }
  h->lag = (int)lag;
  h->ring_slots = lag > 0 ? 2 * (int)lag + 1 : 1;
  for (int i = 0; ok && i < h->nrings; ++i) ok = cudaMalloc(&h->d_ring[i], (size_t)h->ring_slots * tb) == cudaSuccess;
  h->counter_bytes = sizeof(unsigned) * ((size_t)kTsCounterBase + (size_t)kTsMaxStages * h->ring_slots);
  ok = ok && cudaMalloc((void**)&h->d_counters, h->counter_bytes) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { set_error("tiled Stockham plan: device resources", cudaGetLastError()); ts_destroy(h); return nullptr; }
  int n = snprintf(h->name, sizeof(h->name), h->wmode ? "tsw" : "ts");
  for (int i = 0; i < P; ++i) n += snprintf(h->name + n, sizeof(h->name) - n, "%c%d", i ? 'x' : '_', ts_radix(A[i]));
  return h;
}
const char* ts_name(const TsPlanHost* h) { return h->name; }

template <typename T>
int ts_run(TsPlanHost* h, const T* in, T* out, long long batch, int sign, int lm, int sm,
           const cpx<T>* tw, const cpx<T>* twr, cudaStream_t st) {
  if (batch <= 0) return 0;
  TsParams<T> P;
  memset(&P, 0, sizeof(P));
  P.tw = tw; P.twr = twr; P.twR = reinterpret_cast<const cpx<T>*>(h->d_twR);
  for (int i = 0; i < kTsMaxRings; ++i) P.ring[i] = reinterpret_cast<cpx<T>*>(h->d_ring[i]);
  P.counters = h->d_counters;
  P.N = h->N; P.Nc = h->Nc; P.lag = h->lag; P.ring_slots = h->ring_slots;
  P.twR_entries = h->twR_entries;
  ts_build_stages<T>(P, h->Nc, h->P, h->A, h->tw_off, lm, sm, h->wmode);
  const int ns = P.nstages;
  const long long group = P.group_items;

  std::lock_guard<std::mutex> lock(h->mu);
  PF_CUDA_OK(cudaStreamWaitEvent(st, h->done, 0));
  // the prefetch copies 16-byte pieces: the user's input must be aligned to them (rings and tables are)
  const bool with_pre = h->pre && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  auto kern = sign < 0 ? TsKernels<T>::template kern<-1>(with_pre) : TsKernels<T>::template kern<+1>(with_pre);
  // (the pipeline lag and the rings were sized for h->grid; the variant without prefetch may hold fewer or more CTAs,
  //  never more than its own co-resident maximum)
  const int grid_cap = h->grid_v[with_pre ? 1 : 0] < h->grid ? h->grid_v[with_pre ? 1 : 0] : h->grid;
  const size_t smem = h->smem_v[with_pre ? 1 : 0];
  // work-item numbers are 32-bit: very long batches go in several launches
  long long max_tiles = 1;
  for (int i = 0; i < ns; ++i) if (P.st[i].tiles > max_tiles) max_tiles = P.st[i].tiles;
  const long long max_batch = h->wmode
      ? (long long)((0xFFFFFFFFull - 8ull * (unsigned long long)h->grid) / (unsigned long long)max_tiles)      // batch * tiles + workers < 2^32
      : (long long)((0xFFFFFFFFull - 8ull * (unsigned long long)h->grid) / (unsigned long long)group) - (long long)(ns - 1) * h->lag;
  for (long long b0 = 0; b0 < batch; b0 += max_batch) {
    const long long nb = batch - b0 < max_batch ? batch - b0 : max_batch;
    P.in = in + b0 * 2LL * h->Nc; P.out = out + b0 * 2LL * h->Nc; P.batch = nb;
    PF_CUDA_OK(cudaMemsetAsync(h->d_counters, 0, h->counter_bytes, st));
    if constexpr (sizeof(T) == 4) {
      if (h->wmode) {                                             // stage-specialised warps: every stage needs a worker
        const long long total = nb * group;
        P.total_items = (unsigned)(total > 0xFFFFFFFFll ? 0xFFFFFFFFll : total);   // (informational)
        long long g = (total + kTswWarps - 1) / kTswWarps;
        if (g > h->grid) g = h->grid;
        if (g * kTswWarps < ns) g = (ns + kTswWarps - 1) / kTswWarps;
        auto wk = sign < 0 ? TswKernels::kern<-1>() : TswKernels::kern<+1>();
        wk<<<(int)g, kTswWarps * 32, h->smem, st>>>(P, h->twL_entries);
        count_launch();
        PF_CUDA_OK(cudaGetLastError());
        continue;
      }
    }
    const long long total = (nb + (long long)(ns - 1) * h->lag) * group;      // interleaved order incl. pipeline fill / drain slots
    P.total_items = (unsigned)total;
    const long long g = total < grid_cap ? total : grid_cap;
    kern<<<(int)g, kTsThreads, smem, st>>>(P);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
  }
  PF_CUDA_OK(cudaEventRecord(h->done, st));
  return 0;
}
template int ts_run<float>(TsPlanHost*, const float*, float*, long long, int, int, int, const cpx<float>*, const cpx<float>*, cudaStream_t);
template int ts_run<double>(TsPlanHost*, const double*, double*, long long, int, int, int, const cpx<double>*, const cpx<double>*, cudaStream_t);

}  // namespace pf
