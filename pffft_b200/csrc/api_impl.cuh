// api_impl.cuh -- stamps out the extern "C" surface for one precision.
// PF_API(pffft_, pffftb_, PFFFT_Setup, float, Hooks) defines exactly the symbols declared in
// include/pffft/pffft.h (+ the batched ones of pffft_b200.h); the double header likewise.
#pragma once
#include "engine.cuh"

#define PF_CAT_(a, b) a##b
#define PF_CAT(a, b) PF_CAT_(a, b)

#define PF_API(PFX, BPFX, SETUP_T, T, HOOKS, PERNAME)                                                          \
  struct SETUP_T : public pf::Setup<T> {};                                                                     \
  extern "C" {                                                                                                 \
  PFFFT_EXPORT SETUP_T* PF_CAT(PFX, new_setup)(int N, pffft_transform_t tr) {                                  \
    return pf::engine_new_setup<T, HOOKS, SETUP_T>(N, (int)tr);                                                \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PFX, destroy_setup)(SETUP_T* s) { pf::engine_destroy_setup<T, SETUP_T>(s); }        \
  PFFFT_EXPORT void PF_CAT(PFX, transform)(SETUP_T* s, const T* in, T* out, T* /*work*/, pffft_direction_t d) {\
    pf::engine_transform<T, HOOKS>(s, in, out, 1, (int)d, 0);                                                  \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PFX, transform_ordered)(SETUP_T* s, const T* in, T* out, T* /*work*/,               \
                                                   pffft_direction_t d) {                                      \
    pf::engine_transform<T, HOOKS>(s, in, out, 1, (int)d, 1);                                                  \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PFX, zreorder)(SETUP_T* s, const T* in, T* out, pffft_direction_t d) {              \
    pf::engine_zreorder<T>(s, in, out, 1, (int)d);                                                             \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PFX, zconvolve_accumulate)(SETUP_T* s, const T* a, const T* b, T* ab, T sc) {       \
    pf::engine_zconvolve<T>(s, a, b, ab, sc, 1, 0, 1);                                                         \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PFX, zconvolve_no_accu)(SETUP_T* s, const T* a, const T* b, T* ab, T sc) {          \
    pf::engine_zconvolve<T>(s, a, b, ab, sc, 1, 0, 0);                                                         \
  }                                                                                                            \
  PFFFT_EXPORT int PF_CAT(BPFX, transform_batch)(SETUP_T* s, const T* in, T* out, size_t batch,                \
                                                 pffft_direction_t d, int ordered) {                           \
    return pf::engine_transform<T, HOOKS>(s, in, out, (long long)batch, (int)d, ordered ? 1 : 0);              \
  }                                                                                                            \
  PFFFT_EXPORT int PF_CAT(BPFX, zreorder_batch)(SETUP_T* s, const T* in, T* out, size_t batch,                 \
                                                pffft_direction_t d) {                                         \
    return pf::engine_zreorder<T>(s, in, out, (long long)batch, (int)d);                                       \
  }                                                                                                            \
  PFFFT_EXPORT int PF_CAT(BPFX, zconvolve_batch)(SETUP_T* s, const T* a, const T* b, T* ab, T sc,              \
                                                 size_t batch, int b_shared, int acc) {                        \
    return pf::engine_zconvolve<T>(s, a, b, ab, sc, (long long)batch, b_shared, acc);                          \
  }                                                                                                            \
  PFFFT_EXPORT size_t PF_CAT(BPFX, PERNAME)(const SETUP_T* s) { return s ? s->per() : 0; }                     \
  PFFFT_EXPORT const char* PF_CAT(BPFX, setup_kernel)(const SETUP_T* s) { return s ? s->kernel_name : ""; }    \
  PFFFT_EXPORT int PF_CAT(BPFX, set_stream)(SETUP_T* s, void* st) {                                            \
    if (!s) return (int)cudaErrorInvalidValue;                                                                 \
    s->stream = (cudaStream_t)st; return 0;                                                                    \
  }                                                                                                            \
  /* self-test hooks the reference's tests link against (tests/test_pffft.c:269-272; ref impl            \
     pffft_priv_impl.h:1830,1889 validates its SIMD macros).  Here: known-answer transforms on the device. */ \
  PFFFT_EXPORT int PF_CAT(PF_CAT(validate_, PFX), simd_ex)(FILE* dbg) {                                        \
    return pf::engine_selftest<T, HOOKS, SETUP_T>(dbg);                                                        \
  }                                                                                                            \
  PFFFT_EXPORT void PF_CAT(PF_CAT(validate_, PFX), simd)(void) {                                               \
    pf::engine_selftest<T, HOOKS, SETUP_T>(stdout);                                                            \
  }                                                                                                            \
  PFFFT_EXPORT int PF_CAT(BPFX, setup_tables)(SETUP_T* s, void** p, size_t* n) {                               \
    if (!s || !p || !n) return (int)cudaErrorInvalidValue;                                                     \
    *p = s->d_tables; *n = s->table_bytes; return 0;                                                           \
  }                                                                                                            \
  }
