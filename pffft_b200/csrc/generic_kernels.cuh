// generic_kernels.cuh -- kernels that work for EVERY size the API admits
// (N = 16*2^a*3^b*5^c complex, 32*... real; ref pffft_priv_impl.h:91-98), both precisions.
//
//   k_smem_fft   : one CTA keeps whole transforms in shared memory: load+pre-rotation,
//                  all Stockham radix-{2,3,4,5} stages, post-rotation+store.  One HBM read and
//                  one HBM write per transform.  Used when 2 buffers of Nc complex fit in SMEM.
//   k_glob_*     : the same stages with the ping-pong buffers in global memory, one launch per
//                  stage, for sizes beyond shared memory (up to the reference's 2^26 limit,
//                  pffft_priv_impl.h:1069).
//   k_zreorder   : pffft_zreorder            (pffft_priv_impl.h:1158-1193)
//   k_zconvolve  : pffft_zconvolve_accumulate/_no_accu (pffft_priv_impl.h:1534-1684)
//
// Real transforms are an Nc = N/2 point complex FFT on packed pairs plus a rotation
// (SURVEY App. F); this replaces the reference's rfftf1/rfftb1 + real_finalize/preprocess
// (pffft_priv_impl.h:809-901, :1273-1462) while producing the same canonical / z-domain layouts.
//
// Size-tuned kernels for the headline sizes live in fast_kernels.cuh; these are the complete,
// always-available path and the semantic definition the fast kernels are tested against.
#pragma once
#include "butterfly.cuh"
#include "layout.cuh"

namespace pf {

enum LoadMode  { L_C_ORD = 0, L_C_Z = 1, L_R_TIME = 2, L_R_ORD = 3, L_R_Z = 4 };
enum StoreMode { S_C_ORD = 0, S_C_Z = 1, S_R_TIME = 2, S_R_ORD = 3, S_R_Z = 4 };

#define PF_MAX_FACTORS 28

template <typename T> struct XformParams {
  const T* in;
  T* out;
  long long in_stride;    // elements between consecutive transforms of the batch (input side)
  long long out_stride;   // same, output side
  long long in_limit;     // L_R_TIME only: number of readable elements starting at `in` (<0: no limit);
                          // samples beyond it read as zero (overlap-save tail padding, ref pffastconv.c:231-233)
  int out_count;          // S_R_TIME only: leading real samples stored per transform (ref pffastconv.c:255)
  int in_group;           // transforms per input group (1 = plain batch): transform tr reads from
  int in_gstep;           //   in + (tr / in_group) * in_stride + (tr % in_group) * in_gstep   (split large-N rows)
  int in_estride;         // L_C_ORD only: distance (in complex elements) between consecutive input points (1 = dense);
                          // >1 feeds the decimated sub-sequences of the split large-N path
  long long batch;
  int N;                  // transform length as the API sees it
  int Nc;                 // complex core length: N (complex) or N/2 (real)
  int nfac;
  int fac[PF_MAX_FACTORS];
  unsigned magic_nc;                // ceil(2^32 / Nc)
  unsigned magic[PF_MAX_FACTORS];   // per stage: ceil(2^32 / s) for s = product of the earlier radices (0 when s == 1)
  const cpx<T>* tw;       // exp(-2 pi i k / Nc), k < Nc
  const cpx<T>* twr;      // exp(-2 pi i k / N),  k < N/2   (real transforms)
};

// ------------------------------------------------------------------ element-wise load / store
template <bool ZLAYOUT, bool REAL, bool SWZ = false, typename T> PF_HD cpx<T> spec_get(const T* base, int k, int N) {
  if (!ZLAYOUT) return reinterpret_cast<const cpx<T>*>(base)[k];
  const int p = zpos<REAL>(k, N);
  if (SWZ) return mk<T>(base[zswz(p)], base[zswz(p + 4)]);
  return mk<T>(base[p], base[p + 4]);
}
template <bool ZLAYOUT, bool REAL, bool SWZ = false, typename T> PF_HD void spec_put(T* base, int k, int N, cpx<T> v) {
  if (!ZLAYOUT) { reinterpret_cast<cpx<T>*>(base)[k] = v; return; }
  const int p = zpos<REAL>(k, N);
  if (SWZ) { base[zswz(p)] = v.x; base[zswz(p + 4)] = v.y; return; }
  base[p] = v.x; base[p + 4] = v.y;
}

// element i of the complex core's INPUT for this transform
template <int LM, typename T, bool SWZ = false>
PF_HD cpx<T> load_core(const T* base, int i, int N, int Nc, const cpx<T>* twr, long long avail, bool vec_ok, int es = 1) {
  if (LM == L_C_ORD) return reinterpret_cast<const cpx<T>*>(base)[(long long)i * es];
  if (LM == L_C_Z)   return spec_get<true, false, SWZ>(base, i, N);
  if (LM == L_R_TIME) {
    const long long e = 2LL * i;
    if (vec_ok && (avail < 0 || e + 1 < avail)) return reinterpret_cast<const cpx<T>*>(base)[i];
    const T x = (avail < 0 || e < avail) ? base[e] : T(0);
    const T y = (avail < 0 || e + 1 < avail) ? base[e + 1] : T(0);
    return mk<T>(x, y);
  }
  // backward real: rebuild the packed half-length spectrum Z' = 2*(E + i O) from X (SURVEY App. F;
  // the factor 2 keeps BACKWARD(FORWARD(x)) = N x, ref pffft.h:134)
  constexpr bool Z = (LM == L_R_Z);
  if (i == 0) {
    const cpx<T> s0 = spec_get<Z, true, SWZ>(base, 0, N);      // (X[0], X[N/2])
    return mk<T>(s0.x + s0.y, s0.x - s0.y);
  }
  const cpx<T> a = spec_get<Z, true, SWZ>(base, i, N);
  const cpx<T> b = conj(spec_get<Z, true, SWZ>(base, Nc - i, N));
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul_dir<+1>(d, twr[i]);               // d * exp(+2 pi i k/N)
  return mk<T>(s.x - u.y, s.y + u.x);                     // s + i u
}

// store element k of the transform's OUTPUT given the complex core's natural-order result z[0..Nc)
// CG: z was written by other SMs during this launch (ring buffers of the tiled Stockham pipeline): read it through L2 only
template <bool CG, typename T> PF_HD cpx<T> core_get(const cpx<T>* p) {
#ifdef __CUDA_ARCH__
  if constexpr (CG) {
    if constexpr (sizeof(T) == 4) { const float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return mk<T>(v.x, v.y); }
    else { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return mk<T>(v.x, v.y); }
  }
#endif
  return *p;
}
template <int SM, typename T, bool CG = false>
PF_HD void store_core(T* base, const cpx<T>* z, int k, int N, int Nc, const cpx<T>* twr, int out_count, bool vec_ok) {
  if (SM == S_C_ORD) { spec_put<false, false>(base, k, N, core_get<CG>(z + k)); return; }
  if (SM == S_C_Z)   { spec_put<true, false>(base, k, N, core_get<CG>(z + k)); return; }
  if (SM == S_R_TIME) {
    const cpx<T> v = core_get<CG>(z + k);
    const int e = 2 * k;
    if (vec_ok && e + 1 < out_count) { reinterpret_cast<cpx<T>*>(base)[k] = v; return; }
    if (e < out_count) base[e] = v.x;
    if (e + 1 < out_count) base[e + 1] = v.y;
    return;
  }
  // forward real: X[k] = (Z[k] + conj Z[M-k])/2 - (i/2) W^k (Z[k] - conj Z[M-k])
  constexpr bool Z = (SM == S_R_Z);
  if (k == 0) {
    const cpx<T> z0 = core_get<CG>(z);
    spec_put<Z, true>(base, 0, N, mk<T>(z0.x + z0.y, z0.x - z0.y));
    return;
  }
  const cpx<T> a = core_get<CG>(z + k);
  const cpx<T> b = conj(core_get<CG>(z + (Nc - k)));
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul(d, twr[k]);
  spec_put<Z, true>(base, k, N, scale(s + mul_mi(u), T(0.5)));
}

// ------------------------------------------------------------------ one Stockham butterfly
// Autosort DIF stage: butterfly b of Nc/R; `s` = product of the radices already applied.
// Reads x[b + j*Nc/R] (unit stride in b -> conflict-free / coalesced), writes
// y[q + s*(R*p + k)] * W_Nc^{s*p*k} with p = b/s, q = b%s.
// b / s without a division: s is a per-stage constant, magic = ceil(2^32/s) (exact for b*s < 2^32)
PF_HD int div_by_stage(int b, int s, unsigned magic) {
  if (s == 1) return b;
#ifdef __CUDA_ARCH__
  return (int)__umulhi((unsigned)b, magic);
#else
  return (int)(((unsigned long long)(unsigned)b * magic) >> 32);
#endif
}
inline unsigned stage_magic(int s) { return s == 1 ? 0u : (unsigned)(0xFFFFFFFFull / (unsigned)s + 1ull); }

// core of one butterfly given sp = s * (b / s)
template <int R, int SIGN, typename T>
PF_HD void stockham_bfly_sp(const cpx<T>* x, cpx<T>* y, int b, int Nc, int s, int sp, const cpx<T>* tw) {
  const int m = Nc / R;
  const int q = b - sp;
  cpx<T> a[R];
#pragma unroll
  for (int j = 0; j < R; ++j) a[j] = x[b + j * m];
  dftR<R, SIGN>(a);
  const int o = q + sp * R;
  y[o] = a[0];
  if (sp == 0) {
#pragma unroll
    for (int k = 1; k < R; ++k) y[o + s * k] = a[k];
  } else {
#pragma unroll
    for (int k = 1; k < R; ++k) y[o + s * k] = cmul_dir<SIGN>(a[k], tw[sp * k]);
  }
}
template <int R, int SIGN, typename T>
PF_HD void stockham_bfly(const cpx<T>* x, cpx<T>* y, int b, int Nc, int s, unsigned magic, const cpx<T>* tw) {
  stockham_bfly_sp<R, SIGN>(x, y, b, Nc, s, div_by_stage(b, s, magic) * s, tw);   // shared-memory sizes: b*s < 2^32 holds
}
// global-memory stages (Nc up to 2^26): the 32-bit reciprocal is NOT exact there once s has a factor 3 or 5
// (e.g. Nc = 589824, s = 196608, b = 196607), so these take the exact quotient
template <int SIGN, typename T>
PF_HD void stockham_any(int r, const cpx<T>* x, cpx<T>* y, int b, int Nc, int s, const cpx<T>* tw) {
  const int sp = (b / s) * s;
  switch (r) {
    case 2: stockham_bfly_sp<2, SIGN>(x, y, b, Nc, s, sp, tw); break;
    case 3: stockham_bfly_sp<3, SIGN>(x, y, b, Nc, s, sp, tw); break;
    case 4: stockham_bfly_sp<4, SIGN>(x, y, b, Nc, s, sp, tw); break;
    default: stockham_bfly_sp<5, SIGN>(x, y, b, Nc, s, sp, tw); break;
  }
}
// one whole stage of one transform for the threads li, li+tpt, ... (radix dispatch hoisted out of the loop)
template <int R, int SIGN, typename T>
PF_HD void stockham_stage_r(const cpx<T>* x, cpx<T>* y, int li, int tpt, int Nc, int s, unsigned magic, const cpx<T>* tw) {
  const int m = Nc / R;
  for (int b = li; b < m; b += tpt) stockham_bfly<R, SIGN>(x, y, b, Nc, s, magic, tw);
}
template <int SIGN, typename T>
PF_HD void stockham_stage(int r, const cpx<T>* x, cpx<T>* y, int li, int tpt, int Nc, int s, unsigned magic, const cpx<T>* tw) {
  switch (r) {
    case 2: stockham_stage_r<2, SIGN>(x, y, li, tpt, Nc, s, magic, tw); break;
    case 3: stockham_stage_r<3, SIGN>(x, y, li, tpt, Nc, s, magic, tw); break;
    case 4: stockham_stage_r<4, SIGN>(x, y, li, tpt, Nc, s, magic, tw); break;
    default: stockham_stage_r<5, SIGN>(x, y, li, tpt, Nc, s, magic, tw); break;
  }
}

template <typename T> PF_HD bool vec_aligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & (2 * sizeof(T) - 1)) == 0; }

#ifdef __CUDACC__
// ------------------------------------------------------------------ shared-memory kernel
// `tpc` transforms are resident per CTA iteration (many for small N, one for large N).
template <typename T, int LM, int SM, int SIGN>
__global__ void __launch_bounds__(256) k_smem_fft(const XformParams<T> p, const int tpc, const int log2_tpt) {
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* bufA = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* bufB = bufA + (size_t)tpc * p.Nc;
  const int Nc = p.Nc, N = p.N;
  const int tid = threadIdx.x;
  const int tpt = 1 << log2_tpt;               // threads that share one transform (blockDim.x == tpc * tpt)
  const int tl = tid >> log2_tpt;              // which of the CTA's resident transforms
  const int li = tid & (tpt - 1);              // lane inside that transform's thread group

  for (long long t0 = (long long)blockIdx.x * tpc; t0 < p.batch; t0 += (long long)gridDim.x * tpc) {
    const bool live = t0 + tl < p.batch;
    cpx<T>* A = bufA + (size_t)tl * Nc;
    cpx<T>* B = bufB + (size_t)tl * Nc;
    // ---- load (+ z-domain gather / backward-real pre-rotation).  Flattened over the CTA's resident transforms so
    // that consecutive threads touch consecutive elements even when a transform has only a few threads of its own.
    if (log2_tpt < 3) {                        // only 2-4 threads per transform: flatten (measured: N=16c 0.25 vs 0.16)
      const int nt = (int)((p.batch - t0 < tpc) ? (p.batch - t0) : tpc);
      for (int idx = tid; idx < nt * Nc; idx += blockDim.x) {
        const int t2 = div_by_stage(idx, Nc, p.magic_nc), i = idx - t2 * Nc;
        const T* base = p.in + (t0 + t2) * p.in_stride;
        const long long avail = (p.in_limit < 0) ? -1 : (p.in_limit - (t0 + t2) * p.in_stride);
        bufA[idx] = load_core<LM, T>(base, i, N, Nc, p.twr, avail, vec_aligned<T>(base));
      }
    } else if (live) {                         // whole warps per transform: plain strided loop, no index division
      const T* base = p.in + (t0 + tl) * p.in_stride;
      const long long avail = (p.in_limit < 0) ? -1 : (p.in_limit - (t0 + tl) * p.in_stride);
      const bool vok = vec_aligned<T>(base);
      for (int i = li; i < Nc; i += tpt) A[i] = load_core<LM, T>(base, i, N, Nc, p.twr, avail, vok);
    }
    __syncthreads();
    // ---- Stockham stages, ping-pong A <-> B
    cpx<T>* src = A;
    cpx<T>* dst = B;
    int s = 1;
    for (int f = 0; f < p.nfac; ++f) {
      const int r = p.fac[f];
      if (live) stockham_stage<SIGN, T>(r, src, dst, li, tpt, Nc, s, p.magic[f], p.tw);
      __syncthreads();
      cpx<T>* t = src; src = dst; dst = t;
      s *= r;
    }
    // ---- store (+ forward-real post-rotation / z-domain scatter), flattened like the load
    if (log2_tpt < 3) {
      const int nt = (int)((p.batch - t0 < tpc) ? (p.batch - t0) : tpc);
      const cpx<T>* res = (src == A) ? bufA : bufB;            // same ping-pong parity for every resident transform
      for (int idx = tid; idx < nt * Nc; idx += blockDim.x) {
        const int t2 = div_by_stage(idx, Nc, p.magic_nc), k = idx - t2 * Nc;
        T* base = p.out + (t0 + t2) * p.out_stride;
        store_core<SM, T>(base, res + (size_t)t2 * Nc, k, N, Nc, p.twr, p.out_count, vec_aligned<T>(base));
      }
    } else if (live) {
      T* base = p.out + (t0 + tl) * p.out_stride;
      const bool vok = vec_aligned<T>(base);
      for (int k = li; k < Nc; k += tpt) store_core<SM, T>(base, src, k, N, Nc, p.twr, p.out_count, vok);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ global-memory passes (large N)
template <typename T, int LM>
__global__ void __launch_bounds__(256) k_glob_load(const XformParams<T> p, cpx<T>* __restrict__ dst) {
  const long long total = p.batch * p.Nc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / p.Nc;
    const int i = (int)(idx - t * p.Nc);
    const T* base = p.in + t * p.in_stride;
    const long long avail = (p.in_limit < 0) ? -1 : (p.in_limit - t * p.in_stride);
    dst[idx] = load_core<LM, T>(base, i, p.N, p.Nc, p.twr, avail, vec_aligned<T>(base));
  }
}
template <typename T, int SIGN>
__global__ void __launch_bounds__(256) k_glob_stage(const cpx<T>* __restrict__ src, cpx<T>* __restrict__ dst,
                                                    long long batch, int Nc, int r, int s, const cpx<T>* __restrict__ tw) {
  const int m = Nc / r;
  const long long total = batch * m;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / m;
    const int b = (int)(idx - t * m);
    stockham_any<SIGN, T>(r, src + t * Nc, dst + t * Nc, b, Nc, s, tw);
  }
}
template <typename T, int SM>
__global__ void __launch_bounds__(256) k_glob_store(const XformParams<T> p, const cpx<T>* __restrict__ src) {
  const long long total = p.batch * p.Nc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / p.Nc;
    const int k = (int)(idx - t * p.Nc);
    T* base = p.out + t * p.out_stride;
    store_core<SM, T>(base, src + t * p.Nc, k, p.N, p.Nc, p.twr, p.out_count, vec_aligned<T>(base));
  }
}

// ------------------------------------------------------------------ large N: last radix-R step of the split path
// Nc = R * N2.  The R decimated sub-sequences x[n1 + R*n2] were transformed by the CTA kernel into Y[n1][k2]
// (rows of N2); this kernel finishes with   X[k2 + N2*k1] = sum_n1 W_R^{n1 k1} * (W_Nc^{n1 k2} Y[n1][k2]).
// One thread per (transform, k2): R coalesced reads, R coalesced writes.  tw = exp(-2 pi i k / Nc).
// ROWMAJOR: tw[n1*N2 + k2] (unit-stride reads) instead of the natural table read at n1*k2
template <typename T, int R, int SIGN, bool ROWMAJOR = false>
__global__ void __launch_bounds__(256) k_split_combine(const cpx<T>* __restrict__ Y, cpx<T>* __restrict__ X, long long batch,
                                                       int N2, const cpx<T>* __restrict__ tw) {
  const long long total = batch * N2;
  constexpr int bits = ct::ilog2(R);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / N2;
    const int k2 = (int)(idx - t * N2);
    const cpx<T>* y = Y + t * (long long)R * N2 + k2;
    cpx<T> v[R];
#pragma unroll
    for (int p = 0; p < R; ++p) {
      const int n1 = ct::bitrev(p, bits);
      const cpx<T> a = y[(long long)n1 * N2];
      v[p] = (n1 == 0) ? a : cmul_dir<SIGN>(a, ROWMAJOR ? tw[(long long)n1 * N2 + k2] : tw[(long long)n1 * k2]);
    }
    reg_fft<R, SIGN>(v);
    cpx<T>* x = X + t * (long long)R * N2 + k2;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) x[(long long)k1 * N2] = v[k1];
  }
}

// same, for any radix the register DFT library offers (3,5,6,9,10,12,15 ... natural order in and out)
template <typename T, int R, int SIGN, bool ROWMAJOR = false>
__global__ void __launch_bounds__(256) k_split_combine_any(const cpx<T>* __restrict__ Y, cpx<T>* __restrict__ X, long long batch,
                                                           int N2, const cpx<T>* __restrict__ tw) {
  const long long total = batch * N2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / N2;
    const int k2 = (int)(idx - t * N2);
    const cpx<T>* y = Y + t * (long long)R * N2 + k2;
    cpx<T> v[R];
#pragma unroll
    for (int n1 = 0; n1 < R; ++n1) {
      const cpx<T> a = y[(long long)n1 * N2];
      v[n1] = (n1 == 0) ? a : cmul_dir<SIGN>(a, ROWMAJOR ? tw[(long long)n1 * N2 + k2] : tw[(long long)n1 * k2]);
    }
    dft_small<R, SIGN>(v);
    cpx<T>* x = X + t * (long long)R * N2 + k2;
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) x[(long long)k1 * N2] = v[k1];
  }
}

// 4-element vector accesses (128-bit for float, 2 x 128-bit for double)
template <typename T> struct vec4 { T v[4]; };
template <typename T> PF_D vec4<T> ld4(const T* p);
template <> PF_D vec4<float> ld4<float>(const float* p) { float4 t = *reinterpret_cast<const float4*>(p); return {{t.x, t.y, t.z, t.w}}; }
template <> PF_D vec4<double> ld4<double>(const double* p) {
  double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  return {{a.x, a.y, b.x, b.y}};
}
template <typename T> PF_D void st4(T* p, const vec4<T>& v);
template <> PF_D void st4<float>(float* p, const vec4<float>& v) { *reinterpret_cast<float4*>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]); }
template <> PF_D void st4<double>(double* p, const vec4<double>& v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.v[0], v.v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v.v[2], v.v[3]);
}

// ------------------------------------------------------------------ zreorder (pure permutation)
// ref pffft_priv_impl.h:1158-1193.  TOZ=false: z-domain -> canonical (PFFFT_FORWARD); TOZ=true: canonical -> z-domain
// (PFFFT_BACKWARD).  One thread per z-domain GROUP of 8 elements [4 re | 4 im] = elements u = 4*blk..4*blk+3 of quarter q
// (layout.cuh): consecutive threads take consecutive groups, so the z side is one contiguous 32-byte (float) piece
// per thread and a contiguous 1 KiB per warp, moved with 128-bit accesses.  On the canonical side a group is four
// consecutive complex slots: ascending (complex, and the even quarters of a real spectrum) -> two 128-bit accesses per
// thread, 256 contiguous bytes per quarter per warp; descending (odd quarters of a real spectrum, k = M - u with the
// u = 0 element parked at k = M - N/8) -> four complex-sized accesses.  Every 32-byte sector is used in full both ways.
template <typename T, bool REAL, bool TOZ>
__global__ void __launch_bounds__(256) k_zreorder(const T* __restrict__ in, T* __restrict__ out, long long batch, int N) {
  const long long per = REAL ? N : 2LL * N;
  const int groups = (int)(per >> 3);
  const int nq = REAL ? (N >> 3) : (N >> 2);                 // canonical slots per quarter
  const long long total = batch * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / groups;
    const int g = (int)(idx - t * groups);
    const int blk = g >> 2, q = g & 3;
    const T* ib = in + t * per;
    T* ob = out + t * per;
    const int zoff = 8 * g;                                  // = 32*blk + 8*q
    if (!REAL || !(q & 1)) {
      const int coff = 2 * (q * nq + 4 * blk);               // first element of canonical slot k0 = q*nq + 4*blk
      if (TOZ) {
        const vec4<T> c0 = ld4(ib + coff), c1 = ld4(ib + coff + 4);
        st4(ob + zoff,     vec4<T>{{c0.v[0], c0.v[2], c1.v[0], c1.v[2]}});
        st4(ob + zoff + 4, vec4<T>{{c0.v[1], c0.v[3], c1.v[1], c1.v[3]}});
      } else {
        const vec4<T> zr = ld4(ib + zoff), zi = ld4(ib + zoff + 4);
        st4(ob + coff,     vec4<T>{{zr.v[0], zi.v[0], zr.v[1], zi.v[1]}});
        st4(ob + coff + 4, vec4<T>{{zr.v[2], zi.v[2], zr.v[3], zi.v[3]}});
      }
    } else {
      const int M = (q + 1) * nq;                            // q=1: N/4, q=3: N/2
      int k[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int u = 4 * blk + j; k[j] = (u == 0) ? M - nq : M - u; }
      if (TOZ) {
        vec4<T> zr, zi;
#pragma unroll
        for (int j = 0; j < 4; ++j) { const cpx<T> c = reinterpret_cast<const cpx<T>*>(ib)[k[j]]; zr.v[j] = c.x; zi.v[j] = c.y; }
        st4(ob + zoff, zr);
        st4(ob + zoff + 4, zi);
      } else {
        const vec4<T> zr = ld4(ib + zoff), zi = ld4(ib + zoff + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<cpx<T>*>(ob)[k[j]] = mk<T>(zr.v[j], zi.v[j]);
      }
    }
  }
}

// ------------------------------------------------------------------ zconvolve
// z-domain spectra are groups of 8 elements (4 re, 4 im).  One thread per group: 2x 4-wide
// loads of a, b (and ab when accumulating).  Arithmetic is done with explicitly un-fused
// mul/add in the reference's operation order (VCPLXMUL then VMADD, src/simd/pf_float.h:76,
// pf_sse1_float.h:61) so results are bit-identical to the CPU library; the kernel is
// bandwidth-bound, the extra instructions are free.  Real setups multiply element 0 (DC) and
// element 4 (Nyquist) as independent reals (pffft_priv_impl.h:1626-1629, :1680-1683).
PF_D float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
PF_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
PF_D float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
PF_D double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T, bool REAL, bool ACC>
__global__ void __launch_bounds__(256) k_zconvolve(const T* a, const T* b, T* ab,  /* may alias (ref pffft.h:194,208) */
                                                   T scaling, long long batch, int per /*elements per spectrum*/, int b_shared) {
  const int groups = per / 8;
  const long long total = batch * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / groups;
    const int g = (int)(idx - t * groups);
    const long long off = t * per + 8LL * g;
    const long long boff = (b_shared ? 0 : t * (long long)per) + 8LL * g;
    const vec4<T> ar = ld4(a + off), ai = ld4(a + off + 4);
    const vec4<T> br = ld4(b + boff), bi = ld4(b + boff + 4);
    vec4<T> cr, ci;
    if (ACC) { cr = ld4(ab + off); ci = ld4(ab + off + 4); }
    vec4<T> orr, oi;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const T re = add_rn(mul_rn(ar.v[l], br.v[l]), -mul_rn(ai.v[l], bi.v[l]));
      const T im = add_rn(mul_rn(ai.v[l], br.v[l]), mul_rn(ar.v[l], bi.v[l]));
      if (ACC) { orr.v[l] = add_rn(mul_rn(re, scaling), cr.v[l]); oi.v[l] = add_rn(mul_rn(im, scaling), ci.v[l]); }
      else     { orr.v[l] = mul_rn(re, scaling);                   oi.v[l] = mul_rn(im, scaling); }
    }
    if (REAL && g == 0) {
      const T dc = mul_rn(mul_rn(ar.v[0], br.v[0]), scaling);
      const T ny = mul_rn(mul_rn(ai.v[0], bi.v[0]), scaling);
      orr.v[0] = ACC ? add_rn(cr.v[0], dc) : dc;
      oi.v[0]  = ACC ? add_rn(ci.v[0], ny) : ny;
    }
    st4(ab + off, orr);
    st4(ab + off + 4, oi);
  }
}
#endif  // __CUDACC__

}  // namespace pf
