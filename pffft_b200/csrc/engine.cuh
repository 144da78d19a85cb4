// engine.cuh -- host-side engine behind the C-ABI: plan objects, kernel selection and launch,
// host<->device staging.  Templated on the scalar type; api_float.cu / api_double.cu instantiate it.
//
// Plays the role of pffft_new_setup / pffft_transform_internal in the reference
// (src/pffft_priv_impl.h:1062-1112, :1465-1532) -- validation rules and layouts are the
// reference's, everything else (plan contents, kernels, batching, streams) is this engine's.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "generic_kernels.cuh"
#include "plan.h"
#include "ts.h"

namespace pf {

// ---------------------------------------------------------------- process-wide diagnostics (host_common.cu)
void set_error(const char* where, cudaError_t e);
void set_error_msg(const char* msg);
void count_launch(int n = 1);
bool ptr_is_device(const void* p);   // cudaPointerGetAttributes: device/managed -> true, host/unregistered -> false
bool host_mapped_pointer(const void* p, void** dev);   // page-locked + mapped host memory -> its device alias
bool zero_copy_enabled();

#define PF_CUDA_OK(call)                                        \
  do {                                                          \
    cudaError_t _e = (call);                                    \
    if (_e != cudaSuccess) { ::pf::set_error(#call, _e); return (int)_e; } \
  } while (0)

// One cached int per CUDA device for ONE kernel instantiation.  cudaFuncSetAttribute / occupancy results are per device and
// process-wide (not per thread), so the cache is a plain function-local static indexed by the device ordinal; racing
// first users compute the same value.  Values are stored +1 (0 = not computed on that device yet).
struct PerDeviceInt {
  enum { kMaxDevices = 64 };
  std::atomic<int> v[kMaxDevices];
  // compute() -> value >= 0, or < 0 for "failed, do not cache" (returned as is)
  template <typename F> int get(int dev, F&& compute) {
    if (dev < 0 || dev >= kMaxDevices) return compute();
    const int c = v[dev].load(std::memory_order_acquire);
    if (c) return c - 1;
    const int r = compute();
    if (r >= 0) v[dev].store(r + 1, std::memory_order_release);
    return r;
  }
};
inline int current_device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d; }
// opt a kernel in to `smem` bytes of dynamic shared memory once per device (never lowered afterwards: callers pass the
// largest size the instantiation can need)
template <typename K> int ensure_dyn_smem(PerDeviceInt& flag, int dev, K kern, size_t smem) {
  int rc = 0;
  flag.get(dev, [&]() -> int {
    if (smem > 48 * 1024) rc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return rc ? -1 : 1;
  });
  if (rc) set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", (cudaError_t)rc);
  return rc;
}

enum KernelKind { KK_SMEM = 0, KK_GLOBAL = 1, KK_FAST = 2, KK_SPLIT = 3 };

// per-call options beyond the classic API (used by pffastconv: overlapping strided blocks)
struct XformOpts {
  long long in_stride = -1;    // elements; -1 -> contiguous
  long long out_stride = -1;
  long long in_limit = -1;     // real-time load: readable elements from `in` (zero padding beyond)
  int out_count = -1;          // real-time store: leading samples stored per transform
};

struct Slot {                  // one lane of the host-pointer pipeline
  cudaStream_t stream = nullptr;
  void* d_in = nullptr;
  void* d_out = nullptr;
  void* d_aux = nullptr;       // third operand of host-pointer zconvolve calls (per-batch b)
};

template <typename T> struct Setup {
  int N = 0, transform = 0, Nc = 0;
  int device = 0, sm_count = 0;
  size_t smem_optin = 0;                  // device limit of dynamic shared memory per CTA
  int nfac = 0;
  int fac[PF_MAX_FACTORS];
  // device tables: [tw: Nc cpx][twr: N/2 cpx (real only)][fast-kernel tables]
  void* d_tables = nullptr;
  size_t table_bytes = 0;
  const cpx<T>* tw = nullptr;
  const cpx<T>* twr = nullptr;
  const cpx<T>* tw_fast = nullptr;
  cudaStream_t stream = nullptr;          // device-pointer calls are enqueued here (0 = legacy default stream)
  // kernel choice
  int kind = KK_SMEM;
  int generic_kind = KK_SMEM;             // KK_SMEM / KK_GLOBAL: the always-available path behind a tuned plan's -1
  TsPlanHost* ts = nullptr;               // tiled Stockham pipeline (large cores; ts.cu), owned by the plan
  int fast_variant = 0;
  int split_R = 0, split_N2 = 0;          // Nc = split_R x split_N2 two-level plans (rows on a tuned kernel + radix-R combine)
  bool split_fused = false;               // ... small enough for ONE kernel (rows parked in shared memory): one HBM round trip
  int split_cluster = 0, split_Q = 1;     // ... or ONE kernel on clusters of split_cluster CTAs (rows parked in DSMEM), split_Q rows per CTA
  bool split_t2d = false;                 // ... or the tiled two-dimensional plan (two dense passes)
  bool split_t2d_cluster = false;         //     its cluster-fused form (pass A -> pass C through DSMEM; opt-in)
  void* d_aux_tables = nullptr;           //     tables of the general-radix tiled plan (opt-in; own allocation, freed with the plan)
  int split_mode = 0;                     //     0 strided row reads, 1 rows distributed through DSMEM
  char name_buf[40] = {0};
  int tpc = 1;                            // transforms resident per CTA (shared-memory kernel), a power of two
  int log2_tpt = 8;                       // log2(threads per transform) = log2(256 / tpc)
  size_t smem_bytes = 0;
  const char* kernel_name = "";
  // host-pointer pipeline (3 slots: H2D / kernels / D2H overlap across slots)
  std::mutex mu;
  Slot slot[3];
  size_t slot_elems = 0;                  // capacity of each d_in / d_out / d_aux, in T elements
  void* d_bshared = nullptr;              // one spectrum: the shared operand b of host-pointer zconvolve calls
  cudaEvent_t bshared_ready = nullptr;
  // scratch for the global-memory (large N) path
  cpx<T>* d_scratch[2] = {nullptr, nullptr};
  size_t scratch_cpx = 0;
  std::mutex scratch_mu;                  // enqueue order of scratch users
  cudaEvent_t scratch_done = nullptr;     // ... and their execution order across streams
  int occ_cache[64] = {0};                // resident CTAs/SM of each generic-kernel instantiation (0 = not queried yet)

  size_t per() const { return transform == XF_REAL ? (size_t)N : 2 * (size_t)N; }   // elements per transform
};

template <typename T> XformParams<T> make_params(Setup<T>* s, const T* in, T* out, long long batch, const XformOpts& o);

template <typename T> struct FastHooks {
  // implemented by api_float.cu for the sizes that have tuned kernels; default: none
  static bool plan(Setup<T>*) { return false; }
  static int run(Setup<T>*, const T*, T*, long long, int, int, cudaStream_t, const XformOpts&) { return -1; }
  static size_t extra_table_cpx(int /*N*/, int /*transform*/) { return 0; }
  static void fill_extra_table(int, int, T*) {}
};

// ---------------------------------------------------------------- plan creation / destruction
template <typename T, typename S> void engine_destroy_setup(S* s);

// S = the C-ABI's opaque struct (derives from Setup<T>)
template <typename T, typename Hooks, typename S>
S* engine_new_setup(int N, int transform) {
  if (!pfplan::setup_size_ok(N, transform)) return nullptr;          // ref :1066-1078, :1105-1109
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error_msg("pffft_new_setup: no CUDA device (this engine has no CPU path)"); cudaGetLastError(); return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { set_error_msg("pffft_new_setup: cudaGetDeviceProperties failed"); cudaGetLastError(); return nullptr; }
  if (prop.major < 10) { set_error_msg("pffft_new_setup: device is not sm_100-class; this library contains sm_100a code only"); return nullptr; }

  S* s = new S();
  s->N = N; s->transform = transform; s->Nc = (transform == XF_REAL) ? N / 2 : N;
  s->device = dev; s->sm_count = prop.multiProcessorCount; s->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  std::vector<int> f = pfplan::factorize(s->Nc);
  s->nfac = (int)f.size();
  for (int i = 0; i < s->nfac; ++i) s->fac[i] = f[i];

  // tables
  const size_t n_tw = (size_t)s->Nc, n_twr = (transform == XF_REAL) ? (size_t)N / 2 : 0;
  const size_t n_fast = Hooks::extra_table_cpx(N, transform);
  std::vector<T> host(2 * (n_tw + n_twr + n_fast));
  pfplan::fill_roots<T>(host.data(), (long long)n_tw, s->Nc);
  if (n_twr) pfplan::fill_roots<T>(host.data() + 2 * n_tw, (long long)n_twr, N);
  if (n_fast) Hooks::fill_extra_table(N, transform, host.data() + 2 * (n_tw + n_twr));
  s->table_bytes = host.size() * sizeof(T);
  if (cudaMalloc(&s->d_tables, s->table_bytes) != cudaSuccess ||
      cudaMemcpy(s->d_tables, host.data(), s->table_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("pffft_new_setup: table upload", cudaGetLastError());
    if (s->d_tables) cudaFree(s->d_tables);
    delete s; return nullptr;
  }
  s->tw = reinterpret_cast<const cpx<T>*>(s->d_tables);
  s->twr = n_twr ? s->tw + n_tw : nullptr;
  s->tw_fast = n_fast ? s->tw + n_tw + n_twr : nullptr;

  // kernel family
  const size_t cbytes = sizeof(cpx<T>);
  const size_t smem_cap = (size_t)prop.sharedMemPerBlockOptin;      // 227 KB on sm_100
  if (2 * (size_t)s->Nc * cbytes + 1024 <= smem_cap) {
    s->kind = KK_SMEM;
    int tpc = 1;                                                   // ~16 KB of transforms per CTA, power of two so
    while (tpc < 128 && (size_t)(2 * tpc) * s->Nc <= 2048) tpc *= 2;  // that thread -> (transform, lane) is shift/mask
    s->tpc = tpc;
    s->log2_tpt = 8; for (int v = tpc; v > 1; v >>= 1) --s->log2_tpt;
    s->smem_bytes = 2 * (size_t)tpc * s->Nc * cbytes;
    s->kernel_name = "smem_stockham";
  } else {
    s->kind = KK_GLOBAL;
    s->kernel_name = "global_stockham";
  }
  s->generic_kind = s->kind;
  if (Hooks::plan(s)) s->kind = KK_FAST;

  // one transform's worth of staging so single host-pointer calls never allocate (README.md:269-271 of the reference)
  for (int i = 0; i < 3; ++i) {
    if (cudaStreamCreateWithFlags(&s->slot[i].stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create", cudaGetLastError()); }
  }
  s->slot_elems = s->per();
  bool ok = true;
  for (int i = 0; i < 3 && ok; ++i) {
    ok = cudaMalloc(&s->slot[i].d_in, s->slot_elems * sizeof(T)) == cudaSuccess &&
         cudaMalloc(&s->slot[i].d_out, s->slot_elems * sizeof(T)) == cudaSuccess &&
         cudaMalloc(&s->slot[i].d_aux, s->slot_elems * sizeof(T)) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&s->d_bshared, s->per() * sizeof(T)) == cudaSuccess &&
       cudaEventCreateWithFlags(&s->bshared_ready, cudaEventDisableTiming) == cudaSuccess;
  if (ok && s->kind == KK_GLOBAL) {                                  // (tuned large-N plans allocate it on first fallback use)
    s->scratch_cpx = (size_t)s->Nc;
    ok = cudaMalloc((void**)&s->d_scratch[0], s->scratch_cpx * cbytes) == cudaSuccess &&
         cudaMalloc((void**)&s->d_scratch[1], s->scratch_cpx * cbytes) == cudaSuccess &&
         cudaEventCreateWithFlags(&s->scratch_done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) { set_error("pffft_new_setup: device allocation", cudaGetLastError()); engine_destroy_setup<T, S>(s); return nullptr; }
  return s;
}

template <typename T, typename S> void engine_destroy_setup(S* s) {
  if (!s) return;
  int cur = 0; cudaGetDevice(&cur);
  if (cur != s->device) cudaSetDevice(s->device);
  for (int i = 0; i < 3; ++i) {
    if (s->slot[i].stream) { cudaStreamSynchronize(s->slot[i].stream); cudaStreamDestroy(s->slot[i].stream); }
    if (s->slot[i].d_in) cudaFree(s->slot[i].d_in);
    if (s->slot[i].d_out) cudaFree(s->slot[i].d_out);
    if (s->slot[i].d_aux) cudaFree(s->slot[i].d_aux);
  }
  if (s->d_bshared) cudaFree(s->d_bshared);
  if (s->bshared_ready) cudaEventDestroy(s->bshared_ready);
  for (int i = 0; i < 2; ++i) if (s->d_scratch[i]) cudaFree(s->d_scratch[i]);
  if (s->scratch_done) cudaEventDestroy(s->scratch_done);
  if (s->d_tables) cudaFree(s->d_tables);
  if (s->d_aux_tables) cudaFree(s->d_aux_tables);
  if (s->ts) ts_destroy(s->ts);
  if (cur != s->device) cudaSetDevice(cur);
  delete s;
}

// ---------------------------------------------------------------- scratch shared by the multi-launch paths
// caller holds s->scratch_mu; makes `st` wait for the previous user and grows both buffers to `need` complex words
template <typename T> int scratch_acquire(Setup<T>* s, size_t need, cudaStream_t st) {
  if (!s->scratch_done) PF_CUDA_OK(cudaEventCreateWithFlags(&s->scratch_done, cudaEventDisableTiming));
  PF_CUDA_OK(cudaStreamWaitEvent(st, s->scratch_done, 0));
  if (need > s->scratch_cpx) {
    PF_CUDA_OK(cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) { if (s->d_scratch[i]) cudaFree(s->d_scratch[i]); s->d_scratch[i] = nullptr; }
    PF_CUDA_OK(cudaMalloc((void**)&s->d_scratch[0], need * sizeof(cpx<T>)));
    PF_CUDA_OK(cudaMalloc((void**)&s->d_scratch[1], need * sizeof(cpx<T>)));
    s->scratch_cpx = need;
  }
  return 0;
}

// ---------------------------------------------------------------- launches (device pointers)
template <typename T, int LM, int SM, int SIGN>
int launch_generic(Setup<T>* s, const XformParams<T>& p, cudaStream_t st) {
  if (s->generic_kind == KK_GLOBAL) {
    // the ping-pong scratch is shared by every stream using this plan: serialise its users
    std::lock_guard<std::mutex> lock(s->scratch_mu);
    { const int rc = scratch_acquire(s, (size_t)p.batch * s->Nc, st); if (rc) return rc; }
    const long long total = p.batch * (long long)s->Nc;
    const int thr = 256;
    auto grid_for = [&](long long work) { long long g = (work + thr - 1) / thr; long long cap = (long long)s->sm_count * 32; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); };
    k_glob_load<T, LM><<<grid_for(total), thr, 0, st>>>(p, s->d_scratch[0]);
    count_launch();
    int cur = 0, stride = 1;
    for (int f = 0; f < s->nfac; ++f) {
      const int r = s->fac[f];
      k_glob_stage<T, SIGN><<<grid_for(total / r), thr, 0, st>>>(s->d_scratch[cur], s->d_scratch[cur ^ 1], p.batch, s->Nc, r, stride, s->tw);
      count_launch();
      cur ^= 1; stride *= r;
    }
    k_glob_store<T, SM><<<grid_for(total), thr, 0, st>>>(p, s->d_scratch[cur]);
    count_launch();
    PF_CUDA_OK(cudaGetLastError());
    PF_CUDA_OK(cudaEventRecord(s->scratch_done, st));
    return 0;
  }
  // shared-memory kernel
  auto kern = k_smem_fft<T, LM, SM, SIGN>;
  // plans of different sizes share this instantiation: opt in once per device to the device maximum, never lower it
  static PerDeviceInt attr;
  { const int rc = ensure_dyn_smem(attr, s->device, kern, s->smem_optin); if (rc) return rc; }
  constexpr int combo = (LM * 5 + SM) * 2 + (SIGN > 0 ? 1 : 0);
  int per_sm = s->occ_cache[combo];
  if (per_sm == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, s->smem_bytes);
    if (per_sm < 1) per_sm = 1;
    s->occ_cache[combo] = per_sm;
  }
  long long ctas = (p.batch + s->tpc - 1) / s->tpc;
  const long long cap = (long long)s->sm_count * per_sm;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  kern<<<(int)ctas, 256, s->smem_bytes, st>>>(p, s->tpc, s->log2_tpt);
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T> XformParams<T> make_params(Setup<T>* s, const T* in, T* out, long long batch, const XformOpts& o) {
  XformParams<T> p;
  p.in = in; p.out = out;
  p.in_stride = o.in_stride >= 0 ? o.in_stride : (long long)s->per();
  p.out_stride = o.out_stride >= 0 ? o.out_stride : (long long)s->per();
  p.in_limit = o.in_limit;
  p.out_count = o.out_count >= 0 ? o.out_count : s->N;
  p.in_estride = 1; p.in_group = 1; p.in_gstep = 0;
  p.batch = batch; p.N = s->N; p.Nc = s->Nc; p.nfac = s->nfac;
  p.magic_nc = stage_magic(s->Nc);
  { int prod = 1; for (int i = 0; i < PF_MAX_FACTORS; ++i) { p.fac[i] = i < s->nfac ? s->fac[i] : 1; p.magic[i] = stage_magic(prod); prod *= p.fac[i]; } }
  p.tw = s->tw; p.twr = s->twr;
  return p;
}

// transform `batch` vectors resident on the device
template <typename T, typename Hooks>
int engine_transform_device(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered,
                            cudaStream_t st, const XformOpts& o = XformOpts()) {
  if (batch <= 0) return 0;
  if (s->kind == KK_FAST) {
    const int rc = Hooks::run(s, in, out, batch, direction, ordered, st, o);
    if (rc >= 0) return rc;                       // -1: this (direction, layout, options) has no tuned kernel -> generic
  }
  const XformParams<T> p = make_params(s, in, out, batch, o);
  const bool fwd = direction == DIR_FORWARD;
  if (s->transform == XF_COMPLEX) {
    if (fwd) return ordered ? launch_generic<T, L_C_ORD, S_C_ORD, -1>(s, p, st) : launch_generic<T, L_C_ORD, S_C_Z, -1>(s, p, st);
    return ordered ? launch_generic<T, L_C_ORD, S_C_ORD, +1>(s, p, st) : launch_generic<T, L_C_Z, S_C_ORD, +1>(s, p, st);
  }
  if (fwd) return ordered ? launch_generic<T, L_R_TIME, S_R_ORD, -1>(s, p, st) : launch_generic<T, L_R_TIME, S_R_Z, -1>(s, p, st);
  return ordered ? launch_generic<T, L_R_ORD, S_R_TIME, +1>(s, p, st) : launch_generic<T, L_R_Z, S_R_TIME, +1>(s, p, st);
}

// ---------------------------------------------------------------- host-pointer pipeline
// Streams the batch through three slots (H2D -> kernels -> D2H in slot order on the slot's
// stream); consecutive chunks use different slots so copies in both directions overlap compute.
template <typename T> int ensure_slots(Setup<T>* s, size_t elems) {
  if (elems <= s->slot_elems) return 0;
  for (int i = 0; i < 3; ++i) {
    PF_CUDA_OK(cudaStreamSynchronize(s->slot[i].stream));
    if (s->slot[i].d_in) cudaFree(s->slot[i].d_in);
    if (s->slot[i].d_out) cudaFree(s->slot[i].d_out);
    if (s->slot[i].d_aux) cudaFree(s->slot[i].d_aux);
    s->slot[i].d_in = s->slot[i].d_out = s->slot[i].d_aux = nullptr;
    PF_CUDA_OK(cudaMalloc(&s->slot[i].d_in, elems * sizeof(T)));
    PF_CUDA_OK(cudaMalloc(&s->slot[i].d_out, elems * sizeof(T)));
    PF_CUDA_OK(cudaMalloc(&s->slot[i].d_aux, elems * sizeof(T)));
  }
  s->slot_elems = elems;
  return 0;
}

// makes the plan's device current for the scope of a call and restores the caller's on EVERY exit path
struct DeviceScope {
  int prev = -1; bool switched = false; int rc = 0;
  explicit DeviceScope(int want) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != want) { rc = (int)cudaSetDevice(want); switched = (rc == 0); if (rc) set_error("cudaSetDevice(plan device)", (cudaError_t)rc); }
  }
  ~DeviceScope() { if (switched && prev >= 0) cudaSetDevice(prev); }
};
// waits for everything queued on the slot streams (also on error paths: no copy into or out of a user buffer may still
// be in flight when the call returns)
template <typename T> struct SlotDrain {
  Setup<T>* s; int rc = 0;
  explicit SlotDrain(Setup<T>* s_) : s(s_) {}
  int finish() {
    for (int k = 0; k < 3; ++k) { const cudaError_t e = cudaStreamSynchronize(s->slot[k].stream); if (e != cudaSuccess && !rc) { rc = (int)e; set_error("host pipeline: stream synchronize", e); } }
    s = nullptr;
    return rc;
  }
  ~SlotDrain() { if (s) for (int k = 0; k < 3; ++k) cudaStreamSynchronize(s->slot[k].stream); }
};

template <typename T, typename Fn>
int host_pipeline(Setup<T>* s, const T* in, T* out, long long batch, Fn&& device_op) {
  std::lock_guard<std::mutex> lock(s->mu);
  DeviceScope dev(s->device);
  if (dev.rc) return dev.rc;
  SlotDrain<T> drain(s);
  const size_t per = s->per();
  static const size_t chunk_bytes = []() {                                  // ~32 MiB per direction per chunk (PFFFT_B200_CHUNK_MB overrides)
    const char* e = getenv("PFFFT_B200_CHUNK_MB");
    long mb = e ? atol(e) : 32;
    if (mb < 1) mb = 1; if (mb > 1024) mb = 1024;
    return (size_t)mb << 20;
  }();
  long long chunk = (long long)(chunk_bytes / (per * sizeof(T)));
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  int rc = ensure_slots(s, (size_t)chunk * per);
  if (rc) return rc;
  // (a ramped schedule -- 1 MiB first/last chunks doubling up to 32 MiB to shorten pipeline fill and drain -- was
  //  measured SLOWER on the B200 box: 4.73 M vs 5.12 M FFT/s at 2^16 transforms per call; fixed 32 MiB chunks stay)
  // Output side: when `out` is page-locked AND mapped (pffft_aligned_malloc, cudaHostAlloc, cudaHostRegister) the kernels
  // can store straight into it over PCIe (posted writes, whole 128-byte lines) -- no device staging buffer and no D2H
  // copy; the input keeps coming through the copy engine (SM reads over PCIe are latency bound).  PFFFT_B200_ZC_OUT=1.
  static const bool zc_out = getenv("PFFFT_B200_ZC_OUT") ? atoi(getenv("PFFFT_B200_ZC_OUT")) != 0 : false;
  T* out_alias = nullptr;
  if (zc_out) { void* d = nullptr; if (host_mapped_pointer(out, &d)) out_alias = (T*)d; }
  int i = 0;
  long long b0 = 0;
  auto submit = [&](long long nb) -> int {
    Slot& sl = s->slot[i];
    PF_CUDA_OK(cudaMemcpyAsync(sl.d_in, in + (size_t)b0 * per, (size_t)nb * per * sizeof(T), cudaMemcpyHostToDevice, sl.stream));
    if (out_alias) {
      const int r = device_op((const T*)sl.d_in, out_alias + (size_t)b0 * per, nb, sl.stream);
      if (r) return r;
    } else {
      const int r = device_op((const T*)sl.d_in, (T*)sl.d_out, nb, sl.stream);
      if (r) return r;
      PF_CUDA_OK(cudaMemcpyAsync(out + (size_t)b0 * per, sl.d_out, (size_t)nb * per * sizeof(T), cudaMemcpyDeviceToHost, sl.stream));
    }
    b0 += nb; i = (i + 1) % 3;
    return 0;
  };
  while (b0 < batch) { rc = submit((batch - b0 < chunk) ? (batch - b0) : chunk); if (rc) return rc; }
  return drain.finish();
}

// the public transform: host or device pointers
template <typename T, typename Hooks>
int engine_transform(Setup<T>* s, const T* in, T* out, long long batch, int direction, int ordered) {
  if (!s || !in || !out) { set_error_msg("pffft transform: NULL argument"); return (int)cudaErrorInvalidValue; }
  if (direction != DIR_FORWARD && direction != DIR_BACKWARD) { set_error_msg("pffft transform: bad direction"); return (int)cudaErrorInvalidValue; }
  const bool din = ptr_is_device(in), dout = ptr_is_device(out);
  if (din != dout) { set_error_msg("pffft transform: input and output must both be host or both be device pointers"); return (int)cudaErrorInvalidValue; }
  if (din) {                                                   // enqueue on the PLAN's device whatever the caller's current one is
    DeviceScope dev(s->device);
    if (dev.rc) return dev.rc;
    return engine_transform_device<T, Hooks>(s, in, out, batch, direction, ordered, s->stream);
  }
  // page-locked host buffers (pffft_aligned_malloc, cudaHostAlloc, cudaHostRegister) are mapped into the device's
  // address space: with PFFFT_B200_ZEROCOPY=1 the kernels read and write them directly over PCIe (no staging copies)
  if (zero_copy_enabled()) {
    void *din_p = nullptr, *dout_p = nullptr;
    if (host_mapped_pointer(in, &din_p) && host_mapped_pointer(out, &dout_p)) {
      std::lock_guard<std::mutex> lock(s->mu);
      DeviceScope dev(s->device);
      if (dev.rc) return dev.rc;
      cudaStream_t st = s->slot[0].stream;
      const int rc = engine_transform_device<T, Hooks>(s, (const T*)din_p, (T*)dout_p, batch, direction, ordered, st);
      if (rc) return rc;
      PF_CUDA_OK(cudaStreamSynchronize(st));
      return 0;
    }
  }
  return host_pipeline<T>(s, in, out, batch, [&](const T* di, T* dst_, long long nb, cudaStream_t st) {
    return engine_transform_device<T, Hooks>(s, di, dst_, nb, direction, ordered, st);
  });
}

// ---------------------------------------------------------------- device self-test (validate_pffft_simd_ex)
// known answers: delta -> all ones, constant -> N*delta, for one size per kernel family; returns #failures
template <typename T, typename Hooks, typename S> int engine_selftest(FILE* dbg) {
  int errs = 0;
  const int sizes[4] = {64, 1024, 96, 32768};
  for (int tr = 0; tr < 2; ++tr)
    for (int si = 0; si < 4; ++si) {
      const int N = sizes[si];
      if (!pfplan::setup_size_ok(N, tr)) continue;
      S* s = engine_new_setup<T, Hooks, S>(N, tr);
      if (!s) { ++errs; if (dbg) fprintf(dbg, "selftest: no plan for N=%d\n", N); continue; }
      const size_t per = s->per();
      std::vector<T> x(per, T(0)), y(per, T(-1));
      x[0] = T(1);                                               // unit impulse
      engine_transform<T, Hooks>(s, x.data(), y.data(), 1, DIR_FORWARD, 1);
      double worst = 0;
      if (tr == XF_COMPLEX) { for (int k = 0; k < N; ++k) { worst = fmax(worst, fabs((double)y[2 * k] - 1)); worst = fmax(worst, fabs((double)y[2 * k + 1])); } }
      else { worst = fmax(fabs((double)y[0] - 1), fabs((double)y[1] - 1)); for (int k = 1; k < N / 2; ++k) { worst = fmax(worst, fabs((double)y[2 * k] - 1)); worst = fmax(worst, fabs((double)y[2 * k + 1])); } }
      engine_transform<T, Hooks>(s, y.data(), x.data(), 1, DIR_BACKWARD, 1);   // back: N * impulse
      worst = fmax(worst, fabs((double)x[0] / N - 1));
      for (size_t i = 1; i < per; ++i) worst = fmax(worst, fabs((double)x[i]) / N);
      const double tol = sizeof(T) == 4 ? 1e-5 : 1e-12;
      if (!(worst <= tol)) { ++errs; if (dbg) fprintf(dbg, "selftest FAILED: N=%d %s kernel=%s err=%g\n", N, tr ? "complex" : "real", s->kernel_name, worst); }
      else if (dbg) fprintf(dbg, "selftest ok: N=%d %s kernel=%s err=%g\n", N, tr ? "complex" : "real", s->kernel_name, worst);
      engine_destroy_setup<T, S>(s);
    }
  return errs;
}

// ---------------------------------------------------------------- zreorder / zconvolve
template <typename T> int engine_zreorder_device(Setup<T>* s, const T* in, T* out, long long batch, int direction, cudaStream_t st) {
  const long long groups = batch * (long long)(s->per() / 8);      // one thread per 8-element z-domain group
  long long g = (groups + 255) / 256; const long long cap = (long long)s->sm_count * 32;
  if (g > cap) g = cap; if (g < 1) g = 1;
  const bool toz = direction == DIR_BACKWARD;
  if (s->transform == XF_REAL) {
    if (toz) k_zreorder<T, true, true><<<(int)g, 256, 0, st>>>(in, out, batch, s->N);
    else     k_zreorder<T, true, false><<<(int)g, 256, 0, st>>>(in, out, batch, s->N);
  } else {
    if (toz) k_zreorder<T, false, true><<<(int)g, 256, 0, st>>>(in, out, batch, s->N);
    else     k_zreorder<T, false, false><<<(int)g, 256, 0, st>>>(in, out, batch, s->N);
  }
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T> int engine_zreorder(Setup<T>* s, const T* in, T* out, long long batch, int direction) {
  if (!s || !in || !out) { set_error_msg("pffft_zreorder: NULL argument"); return (int)cudaErrorInvalidValue; }
  if (in == out) { set_error_msg("pffft_zreorder: input and output must not alias (ref pffft_priv_impl.h:1162)"); return (int)cudaErrorInvalidValue; }
  const bool din = ptr_is_device(in), dout = ptr_is_device(out);
  if (din != dout) { set_error_msg("pffft_zreorder: mixed host/device pointers"); return (int)cudaErrorInvalidValue; }
  if (din) { DeviceScope dev(s->device); if (dev.rc) return dev.rc; return engine_zreorder_device<T>(s, in, out, batch, direction, s->stream); }
  return host_pipeline<T>(s, in, out, batch, [&](const T* di, T* dst_, long long nb, cudaStream_t st) {
    return engine_zreorder_device<T>(s, di, dst_, nb, direction, st);
  });
}

template <typename T> int engine_zconvolve_device(Setup<T>* s, const T* a, const T* b, T* ab, T scaling, long long batch,
                                                  int b_shared, int accumulate, cudaStream_t st) {
  const int per = (int)s->per();
  const long long groups = batch * (per / 8);
  long long g = (groups + 255) / 256; const long long cap = (long long)s->sm_count * 32;
  if (g > cap) g = cap; if (g < 1) g = 1;
  const bool real = s->transform == XF_REAL;
  if (real) {
    if (accumulate) k_zconvolve<T, true, true><<<(int)g, 256, 0, st>>>(a, b, ab, scaling, batch, per, b_shared);
    else            k_zconvolve<T, true, false><<<(int)g, 256, 0, st>>>(a, b, ab, scaling, batch, per, b_shared);
  } else {
    if (accumulate) k_zconvolve<T, false, true><<<(int)g, 256, 0, st>>>(a, b, ab, scaling, batch, per, b_shared);
    else            k_zconvolve<T, false, false><<<(int)g, 256, 0, st>>>(a, b, ab, scaling, batch, per, b_shared);
  }
  count_launch();
  PF_CUDA_OK(cudaGetLastError());
  return 0;
}
// host or device pointers; host path stages a, b (and ab when accumulating) through slot memory
template <typename T> int engine_zconvolve(Setup<T>* s, const T* a, const T* b, T* ab, T scaling, long long batch,
                                           int b_shared, int accumulate) {
  if (!s || !a || !b || !ab) { set_error_msg("pffft_zconvolve: NULL argument"); return (int)cudaErrorInvalidValue; }
  if (batch <= 0) return 0;
  const bool da = ptr_is_device(a), db = ptr_is_device(b), dab = ptr_is_device(ab);
  if (da != db || da != dab) { set_error_msg("pffft_zconvolve: mixed host/device pointers"); return (int)cudaErrorInvalidValue; }
  if (da) { DeviceScope dev(s->device); if (dev.rc) return dev.rc; return engine_zconvolve_device<T>(s, a, b, ab, scaling, batch, b_shared, accumulate, s->stream); }
  // host pointers: the same three-slot pipeline as the transforms (a -> d_in, per-batch b -> d_aux, ab <-> d_out), ~32 MiB
  // chunks, no allocation for single-spectrum calls (buffers are sized for one transform when the plan is built;
  // ref README.md:269-271 "the fft functions do not perform any memory allocation")
  std::lock_guard<std::mutex> lock(s->mu);
  DeviceScope dev(s->device);
  if (dev.rc) return dev.rc;
  SlotDrain<T> drain(s);
  const size_t per = s->per();
  long long chunk = (long long)(((size_t)32 << 20) / (per * sizeof(T)));
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  { const int rc = ensure_slots(s, (size_t)chunk * per); if (rc) return rc; }
  if (b_shared) {
    PF_CUDA_OK(cudaMemcpyAsync(s->d_bshared, b, per * sizeof(T), cudaMemcpyHostToDevice, s->slot[0].stream));
    PF_CUDA_OK(cudaEventRecord(s->bshared_ready, s->slot[0].stream));
    for (int k = 1; k < 3; ++k) PF_CUDA_OK(cudaStreamWaitEvent(s->slot[k].stream, s->bshared_ready, 0));
  }
  int i = 0;
  for (long long b0 = 0; b0 < batch; b0 += chunk, i = (i + 1) % 3) {
    const long long nb = batch - b0 < chunk ? batch - b0 : chunk;
    const size_t off = (size_t)b0 * per, bytes = (size_t)nb * per * sizeof(T);
    Slot& sl = s->slot[i];
    PF_CUDA_OK(cudaMemcpyAsync(sl.d_in, a + off, bytes, cudaMemcpyHostToDevice, sl.stream));
    if (!b_shared) PF_CUDA_OK(cudaMemcpyAsync(sl.d_aux, b + off, bytes, cudaMemcpyHostToDevice, sl.stream));
    if (accumulate) PF_CUDA_OK(cudaMemcpyAsync(sl.d_out, ab + off, bytes, cudaMemcpyHostToDevice, sl.stream));
    const T* db_ = b_shared ? (const T*)s->d_bshared : (const T*)sl.d_aux;
    { const int rc = engine_zconvolve_device<T>(s, (const T*)sl.d_in, db_, (T*)sl.d_out, scaling, nb, b_shared, accumulate, sl.stream); if (rc) return rc; }
    PF_CUDA_OK(cudaMemcpyAsync(ab + off, sl.d_out, bytes, cudaMemcpyDeviceToHost, sl.stream));
  }
  return drain.finish();
}

}  // namespace pf
