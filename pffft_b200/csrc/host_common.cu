// host_common.cu -- precision-independent parts of the C-ABI: size algebra, aligned (pinned) host
// memory, diagnostics.  ref: src/pffft_common.c:9-55, src/pffft_priv_impl.h:76-116.
#include <cuda_runtime.h>
#include <atomic>
#include <string>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include "plan.h"
#include "../../include/pffft/pffft_b200.h"

namespace pf {
static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* where, cudaError_t e) {
  g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  cudaGetLastError();   // clear the sticky-less error state
}
void set_error_msg(const char* msg) { g_err = msg; }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
bool ptr_is_device(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
bool host_mapped_pointer(const void* p, void** dev) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
  *dev = a.devicePointer;
  return true;
}
bool zero_copy_enabled() {
  static const bool on = getenv("PFFFT_B200_ZEROCOPY") ? atoi(getenv("PFFFT_B200_ZEROCOPY")) != 0 : false;
  return on;
}
}  // namespace pf

// ---- aligned host memory.  Layout of one allocation: [raw ... | header(16 B: raw ptr, kind) | user (64-B aligned)]
namespace {
enum : uint64_t { kKindMalloc = 0x6d616c6c6f63ull, kKindPinned = 0x70696e6e6564ull };
struct Hdr { void* raw; uint64_t kind; };

void* aligned_alloc_impl(size_t nb) {
  const size_t extra = 64 + sizeof(Hdr);
  void* raw = nullptr;
  uint64_t kind = kKindMalloc;
  // page-locked when a device is present: host-pointer transforms then DMA directly from user buffers
  static int have_dev = -1;
  if (have_dev < 0) { int n = 0; have_dev = (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? 1 : 0; if (!have_dev) cudaGetLastError(); }
  if (have_dev && nb >= 4096 && getenv("PFFFT_B200_NO_PINNED") == nullptr) {
    if (cudaHostAlloc(&raw, nb + extra, cudaHostAllocPortable) == cudaSuccess) kind = kKindPinned;
    else { raw = nullptr; cudaGetLastError(); }
  }
  if (!raw) { raw = malloc(nb + extra); kind = kKindMalloc; }
  if (!raw) return nullptr;
  uintptr_t u = ((uintptr_t)raw + extra) & ~(uintptr_t)63;
  Hdr* h = (Hdr*)(u - sizeof(Hdr));
  h->raw = raw; h->kind = kind;
  return (void*)u;
}
void aligned_free_impl(void* p) {
  if (!p) return;
  Hdr* h = (Hdr*)((uintptr_t)p - sizeof(Hdr));
  if (h->kind == kKindPinned) cudaFreeHost(h->raw); else free(h->raw);
}
}  // namespace

extern "C" {
// float-named and double-named helpers are the same functions in the reference too (pffft_common.c:47-55)
PFFFT_EXPORT void* pffft_aligned_malloc(size_t nb) { return aligned_alloc_impl(nb); }
PFFFT_EXPORT void pffft_aligned_free(void* p) { aligned_free_impl(p); }
PFFFT_EXPORT void* pffftd_aligned_malloc(size_t nb) { return aligned_alloc_impl(nb); }
PFFFT_EXPORT void pffftd_aligned_free(void* p) { aligned_free_impl(p); }
PFFFT_EXPORT int pffft_next_power_of_two(int N) { return pfplan::next_power_of_two(N); }
PFFFT_EXPORT int pffftd_next_power_of_two(int N) { return pfplan::next_power_of_two(N); }
PFFFT_EXPORT int pffft_is_power_of_two(int N) { return pfplan::is_power_of_two(N); }
PFFFT_EXPORT int pffftd_is_power_of_two(int N) { return pfplan::is_power_of_two(N); }

PFFFT_EXPORT int pffft_simd_size(void) { return pfplan::kSimd; }
PFFFT_EXPORT int pffftd_simd_size(void) { return pfplan::kSimd; }
PFFFT_EXPORT const char* pffft_simd_arch(void) { return "sm_100a"; }
PFFFT_EXPORT const char* pffftd_simd_arch(void) { return "sm_100a"; }
PFFFT_EXPORT int pffft_min_fft_size(pffft_transform_t t) { return pfplan::min_fft_size((int)t); }
PFFFT_EXPORT int pffftd_min_fft_size(pffft_transform_t t) { return pfplan::min_fft_size((int)t); }
PFFFT_EXPORT int pffft_is_valid_size(int N, pffft_transform_t t) { return pfplan::is_valid_size(N, (int)t); }
PFFFT_EXPORT int pffftd_is_valid_size(int N, pffft_transform_t t) { return pfplan::is_valid_size(N, (int)t); }
PFFFT_EXPORT int pffft_nearest_transform_size(int N, pffft_transform_t t, int higher) { return pfplan::nearest_transform_size(N, (int)t, higher); }
PFFFT_EXPORT int pffftd_nearest_transform_size(int N, pffft_transform_t t, int higher) { return pfplan::nearest_transform_size(N, (int)t, higher); }

PFFFT_EXPORT const char* pffftb_last_error(void) { return pf::g_err.c_str(); }
PFFFT_EXPORT unsigned long long pffftb_launch_count(void) { return pf::g_launches.load(); }
PFFFT_EXPORT int pffftb_device_synchronize(void) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) pf::set_error("cudaDeviceSynchronize", e);
  return (int)e;
}
}  // extern "C"
