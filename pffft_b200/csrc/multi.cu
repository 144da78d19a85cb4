// multi.cu -- multi-GPU entry points of the C-ABI (include/pffft/pffft_b200.h): batch sharding over the GPUs of
// one node plus ONE NCCL broadcast of the plan tables.  The transforms themselves never communicate (the reference's
// PFFFT_Setup is read-only and every transform independent, include/pffft/pffft.h:102-106), so the only collective is the
// table broadcast that makes every GPU use bit-identical tables.
// NCCL is bound at run time (dlopen of libnccl.so.2: the library has no link-time dependency on it, and inside a PyTorch
// process the already loaded NCCL is the one that gets used).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../../include/pffft/pffft_b200.h"
#include "internal_api.h"

namespace pf {
namespace {

// the slice of nccl.h this file needs (ABI-stable across NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclChar = 0 };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
const Nccl& nccl() {
  static const Nccl n = []() {
    Nccl r;
    if (getenv("PFFFT_B200_NO_NCCL")) return r;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { r.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (r.lib) break; }
    if (!r.lib) return r;
#define PF_SYM(field, sym) *(void**)(&r.field) = dlsym(r.lib, sym)
    PF_SYM(GetUniqueId, "ncclGetUniqueId"); PF_SYM(CommInitRank, "ncclCommInitRank"); PF_SYM(CommInitAll, "ncclCommInitAll");
    PF_SYM(CommDestroy, "ncclCommDestroy"); PF_SYM(Broadcast, "ncclBroadcast"); PF_SYM(GroupStart, "ncclGroupStart");
    PF_SYM(GroupEnd, "ncclGroupEnd"); PF_SYM(GetErrorString, "ncclGetErrorString");
#undef PF_SYM
    r.ok = r.GetUniqueId && r.CommInitRank && r.CommInitAll && r.CommDestroy && r.Broadcast && r.GroupStart && r.GroupEnd;
    return r;
  }();
  return n;
}
int nccl_fail(const char* what, int rc) {
  std::string m = std::string(what) + ": NCCL error " + std::to_string(rc);
  if (nccl().GetErrorString) m += std::string(" (") + nccl().GetErrorString(rc) + ")";
  set_error_msg(m.c_str());
  return (int)cudaErrorUnknown;
}

}  // namespace
}  // namespace pf

struct PFFFTB_Multi {
  int n = 0;
  std::vector<int> dev;
  std::vector<PFFFT_Setup*> setup;
  std::vector<cudaStream_t> stream;       // one stream per device (table broadcast)
  const char* backend = "none";
};

using namespace pf;

extern "C" {

PFFFT_EXPORT void pffftb_multi_destroy(PFFFTB_Multi* m) {
  if (!m) return;
  int cur = 0; cudaGetDevice(&cur);
  for (int g = 0; g < m->n; ++g) {
    cudaSetDevice(m->dev[g]);
    if (g < (int)m->stream.size() && m->stream[g]) { cudaStreamSynchronize(m->stream[g]); cudaStreamDestroy(m->stream[g]); }
    if (g < (int)m->setup.size() && m->setup[g]) pffft_destroy_setup(m->setup[g]);
  }
  cudaSetDevice(cur);
  delete m;
}

PFFFT_EXPORT PFFFTB_Multi* pffftb_multi_new(int N, pffft_transform_t transform, int ngpus) {
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) { cudaGetLastError(); set_error_msg("pffftb_multi_new: no CUDA device"); return nullptr; }
  if (ngpus <= 0 || ngpus > visible) ngpus = visible;
  int cur = 0; cudaGetDevice(&cur);
  PFFFTB_Multi* m = new PFFFTB_Multi();
  m->n = ngpus;
  bool ok = true;
  for (int g = 0; g < ngpus && ok; ++g) {
    m->dev.push_back(g);
    ok = cudaSetDevice(g) == cudaSuccess;
    cudaStream_t st = nullptr;
    ok = ok && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    m->stream.push_back(st);
    PFFFT_Setup* s = ok ? pffft_new_setup(N, transform) : nullptr;     // plan on the now-current device
    m->setup.push_back(s);
    ok = ok && s != nullptr;
  }
  // ---- the one collective of this path: tables of GPU 0 -> every GPU
  if (ok && ngpus > 1) {
    std::vector<void*> tab(ngpus); std::vector<size_t> nb(ngpus);
    for (int g = 0; g < ngpus; ++g) pffftb_setup_tables(m->setup[g], &tab[g], &nb[g]);
    for (int g = 1; g < ngpus; ++g) ok = ok && nb[g] == nb[0];
    const Nccl& nc = nccl();
    bool done = false;
    if (ok && nc.ok) {
      std::vector<ncclComm_t> comm(ngpus, nullptr);
      int rc = nc.CommInitAll(comm.data(), ngpus, m->dev.data());
      if (rc == kNcclSuccess) {
        nc.GroupStart();
        for (int g = 0; g < ngpus && rc == kNcclSuccess; ++g) {
          cudaSetDevice(m->dev[g]);
          rc = nc.Broadcast(tab[g], tab[g], nb[0], kNcclChar, 0, comm[g], m->stream[g]);
        }
        const int rc2 = nc.GroupEnd();
        if (rc == kNcclSuccess) rc = rc2;
        for (int g = 0; g < ngpus; ++g) { cudaSetDevice(m->dev[g]); cudaStreamSynchronize(m->stream[g]); }
        for (int g = 0; g < ngpus; ++g) if (comm[g]) nc.CommDestroy(comm[g]);
        done = rc == kNcclSuccess;
        if (!done) nccl_fail("pffftb_multi_new: ncclBroadcast of the plan tables", rc);
      } else nccl_fail("pffftb_multi_new: ncclCommInitAll", rc);
      if (done) m->backend = "nccl";
    }
    if (ok && !done) {                         // same broadcast without NCCL
      for (int g = 1; g < ngpus && ok; ++g) ok = cudaMemcpyPeer(tab[g], m->dev[g], tab[0], m->dev[0], nb[0]) == cudaSuccess;
      if (ok) m->backend = "memcpy_peer"; else set_error("pffftb_multi_new: cudaMemcpyPeer of the plan tables", cudaGetLastError());
    }
  } else if (ok) m->backend = "single";
  cudaSetDevice(cur);
  if (!ok) { pffftb_multi_destroy(m); return nullptr; }
  return m;
}

PFFFT_EXPORT int pffftb_multi_ngpus(const PFFFTB_Multi* m) { return m ? m->n : 0; }
PFFFT_EXPORT PFFFT_Setup* pffftb_multi_setup(PFFFTB_Multi* m, int gpu) { return (m && gpu >= 0 && gpu < m->n) ? m->setup[gpu] : nullptr; }
PFFFT_EXPORT const char* pffftb_multi_broadcast_backend(const PFFFTB_Multi* m) { return m ? m->backend : ""; }

PFFFT_EXPORT int pffftb_multi_transform_batch(PFFFTB_Multi* m, const float* in, float* out, size_t batch,
                                              pffft_direction_t direction, int ordered) {
  if (!m || !in || !out) { set_error_msg("pffftb_multi_transform_batch: NULL argument"); return (int)cudaErrorInvalidValue; }
  const size_t per = pffftb_floats_per_transform(m->setup[0]);
  std::vector<int> rc(m->n, 0);
  std::vector<std::string> msg(m->n);
  std::vector<std::thread> th;
  for (int g = 0; g < m->n; ++g) {
    const size_t lo = batch * (size_t)g / (size_t)m->n, hi = batch * (size_t)(g + 1) / (size_t)m->n;
    if (hi == lo) continue;
    th.emplace_back([=, &rc, &msg]() {
      cudaSetDevice(m->dev[g]);
      rc[g] = pffftb_transform_batch(m->setup[g], in + lo * per, out + lo * per, hi - lo, direction, ordered);
      if (rc[g]) msg[g] = pffftb_last_error();            // error text is thread-local: carry it to the caller's thread
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < m->n; ++g) if (rc[g]) { set_error_msg(msg[g].c_str()); return rc[g]; }
  return 0;
}

PFFFT_EXPORT int pffftb_multi_transform_shards(PFFFTB_Multi* m, const float* const* in, float* const* out, const size_t* batch,
                                               pffft_direction_t direction, int ordered) {
  if (!m || !in || !out || !batch) { set_error_msg("pffftb_multi_transform_shards: NULL argument"); return (int)cudaErrorInvalidValue; }
  for (int g = 0; g < m->n; ++g) {
    if (!batch[g]) continue;
    const int rc = pffftb_transform_batch(m->setup[g], in[g], out[g], batch[g], direction, ordered);   // switches to the plan's device
    if (rc) return rc;
  }
  return 0;
}

PFFFT_EXPORT int pffftb_multi_synchronize(PFFFTB_Multi* m) {
  if (!m) return (int)cudaErrorInvalidValue;
  int cur = 0; cudaGetDevice(&cur);
  int rc = 0;
  for (int g = 0; g < m->n; ++g) {
    cudaSetDevice(m->dev[g]);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess && !rc) { rc = (int)e; set_error("pffftb_multi_synchronize", e); }
  }
  cudaSetDevice(cur);
  return rc;
}

PFFFT_EXPORT int pffftb_nccl_unique_id(void* id128) {
  const Nccl& nc = nccl();
  if (!id128) return (int)cudaErrorInvalidValue;
  if (!nc.ok) { set_error_msg("pffftb_nccl_unique_id: libnccl.so.2 not available"); return (int)cudaErrorNotSupported; }
  ncclUniqueId id;
  const int rc = nc.GetUniqueId(&id);
  if (rc != kNcclSuccess) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, id.internal, 128);
  return 0;
}

PFFFT_EXPORT int pffftb_setup_broadcast_tables(PFFFT_Setup* s, const void* id128, int rank, int nranks) {
  if (!s || !id128 || rank < 0 || rank >= nranks) { set_error_msg("pffftb_setup_broadcast_tables: bad argument"); return (int)cudaErrorInvalidValue; }
  if (nranks == 1 && !getenv("PFFFT_B200_FORCE_NCCL_1RANK")) return 0;   // (the env switch lets a 1-GPU box run the NCCL calls)
  const Nccl& nc = nccl();
  if (!nc.ok) { set_error_msg("pffftb_setup_broadcast_tables: libnccl.so.2 not available"); return (int)cudaErrorNotSupported; }
  DeviceScope dev(pffftb_setup_device(s));
  if (dev.rc) return dev.rc;
  void* tab = nullptr; size_t nb = 0;
  pffftb_setup_tables(s, &tab, &nb);
  ncclUniqueId id; memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  int rc = nc.CommInitRank(&comm, nranks, id, rank);
  if (rc != kNcclSuccess) return nccl_fail("ncclCommInitRank", rc);
  cudaStream_t st = nullptr;
  PF_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  rc = nc.Broadcast(tab, tab, nb, kNcclChar, 0, comm, st);
  const cudaError_t e = cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  nc.CommDestroy(comm);
  if (rc != kNcclSuccess) return nccl_fail("ncclBroadcast of the plan tables", rc);
  if (e != cudaSuccess) { set_error("pffftb_setup_broadcast_tables", e); return (int)e; }
  return 0;
}

}  // extern "C"
