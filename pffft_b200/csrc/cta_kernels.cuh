// cta_kernels.cuh -- one transform per CTA for complex cores of 512..4096 points (real N up to 8192).
//
// Nc = 16 x 16 x C (C = 2,4,8,16), T = Nc/16 threads, every thread owns 16 points in registers.
// Three register-FFT passes (radix 16, 16, C) over the three index digits
//      n = n_c + C*n_b + 16*C*n_a        k = k_a + 16*k_b + 256*k_c
//   pass 1: thread m=(n_b,n_c): FFT over n_a of x[m + 16C*n_a] (coalesced global loads), * W_Nc^{m k_a}
//   pass 2: thread (k_a,n_c)  : FFT over n_b, * W_{16C}^{n_c k_b}           (in place in shared memory)
//   pass 3: thread (k_a,k_b)  : FFT over n_c -> X[k_a + 16 k_b + 256 k_c]   (coalesced global stores)
// The exchange tile is exactly Nc complex words with an XOR swizzle of the low 4 address bits by
// rotr4(k_a, log2(16/C)); every 64-bit shared access of every pass is bank-conflict free for all C
// (derivation in DESIGN.md section 4).  One HBM read + one HBM write per transform.
//
// Load/store reuse the element functions of generic_kernels.cuh, so the same kernel serves complex
// and real transforms (N/2-point packing + rotation), canonical and z-domain layouts, and the strided /
// zero-padded / truncated accesses of pffastconv.  Replaces, for these sizes, cfftf1/rfftf1/rfftb1 +
// finalize/preprocess + zreorder of the reference (src/pffft_priv_impl.h:809-1048, :1195-1462, :1158-1193).
#pragma once
#include "butterfly.cuh"
#include "generic_kernels.cuh"
#include "fast_kernels.cuh"   // mbarrier / bulk-copy helpers

namespace pf {

template <int C> struct K2 {
  static constexpr int NC = 256 * C;          // complex core length
  static constexpr int T = 16 * C;            // threads per transform
  static constexpr int BC = 16 * C;           // stride of the first digit
  static constexpr int SH = (C == 16) ? 0 : (C == 8) ? 1 : (C == 4) ? 2 : 3;   // log2(16/C)
  PF_HD static int swz(int k_a) { return ((k_a >> SH) | (k_a << (4 - SH))) & 15; }
  PF_HD static int idx(int k_a, int j_b, int j_c) { return k_a * (16 * C) + ((j_b * C + j_c) ^ swz(k_a)); }
};

PF_HD constexpr int brev4(int p) { return ct::bitrev(p, 4); }

// table read through the read-only path (ld.global.nc) on the device
template <typename T> PF_HD cpx<T> ldtab(const cpx<T>* p) {
#ifdef __CUDA_ARCH__
  if constexpr (sizeof(T) == 4) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<T>(v.x, v.y); }
  else { const double2 v = __ldg(reinterpret_cast<const double2*>(p)); return mk<T>(v.x, v.y); }
#else
  return *p;
#endif
}
template <int C> PF_HD constexpr int brevC(int p) { return ct::bitrev(p, ct::ilog2(C)); }

// direct store of output element k from a register (modes that need no partner element)
template <int SM, typename T, bool SWZ = false>
PF_HD void store_elem(T* base, int k, cpx<T> v, int N, int out_count, bool vec_ok) {
  if (SM == S_C_ORD) { spec_put<false, false>(base, k, N, v); return; }
  if (SM == S_C_Z)   { spec_put<true, false, SWZ>(base, k, N, v); return; }
  // S_R_TIME: two real samples, truncated to out_count (pffastconv keeps only the valid ones)
  const int e = 2 * k;
  if (vec_ok && e + 1 < out_count) { reinterpret_cast<cpx<T>*>(base)[k] = v; return; }
  if (e < out_count) base[e] = v.x;
  if (e + 1 < out_count) base[e + 1] = v.y;
}

// ---- pass 1: thread m in [0, 16C)
template <int C, int LM, int SIGN, bool FAST, typename T, bool SWZ = false>
PF_HD void k2_pass1(int m, const T* base, int N, const cpx<T>* twr, long long avail, bool vec_ok,
                    const cpx<T>* tw1, cpx<T>* tile, int es = 1) {
  using K = K2<C>;
  cpx<T> v[16];
  if (FAST && (LM == L_R_TIME || LM == L_C_ORD)) {      // contiguous, aligned, fully in range: plain 64-bit loads
    const cpx<T>* src = reinterpret_cast<const cpx<T>*>(base) + m;
#pragma unroll
    for (int p = 0; p < 16; ++p) v[p] = src[K::BC * brev4(p)];
  } else {
#pragma unroll
    for (int p = 0; p < 16; ++p) v[p] = load_core<LM, T, SWZ>(base, m + K::BC * brev4(p), N, K::NC, twr, avail, vec_ok, es);
  }
  reg_fft<16, SIGN>(v);
  const int jb = m / C, jc = m % C;
  tile[K::idx(0, jb, jc)] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[K::idx(ka, jb, jc)] = cmul_dir<SIGN>(v[ka], ldtab(tw1 + ka * K::BC + m));
}
// ---- pass 2: thread t -> (k_a = t / C, n_c = t % C), in place
template <int C, int SIGN, typename T>
PF_HD void k2_pass2(int t, const cpx<T>* tw2, cpx<T>* tile) {
  using K = K2<C>;
  const int ka = t / C, nc = t % C;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = tile[K::idx(ka, brev4(p), nc)];
  reg_fft<16, SIGN>(v);
  tile[K::idx(ka, 0, nc)] = v[0];
#pragma unroll
  for (int kb = 1; kb < 16; ++kb) tile[K::idx(ka, kb, nc)] = cmul_dir<SIGN>(v[kb], ldtab(tw2 + kb * C + nc));
}
// ---- pass 3: thread t -> k_a = t % 16, k_b = t / 16 + C*r (r < 16/C); u[r*C + k_c] = X[k_a + 16 k_b + 256 k_c]
template <int C, int SIGN, typename T>
PF_HD void k2_pass3(int t, const cpx<T>* tile, cpx<T> (&u)[16]) {
  using K = K2<C>;
  const int ka = t & 15, kb0 = t >> 4;
#pragma unroll
  for (int r = 0; r < 16 / C; ++r) {
#pragma unroll
    for (int p = 0; p < C; ++p) u[r * C + p] = tile[K::idx(ka, kb0 + C * r, brevC<C>(p))];
  }
  if (C == 16) dit_fft<16, SIGN, 0, 1>(u);
  if (C == 8) { dit_fft<8, SIGN, 0, 1>(u); dit_fft<8, SIGN, 8, 1>(u); }
  if (C == 4) { dit_fft<4, SIGN, 0, 1>(u); dit_fft<4, SIGN, 4, 1>(u); dit_fft<4, SIGN, 8, 1>(u); dit_fft<4, SIGN, 12, 1>(u); }
  if (C == 2) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { const cpx<T> a = u[2 * r], b = u[2 * r + 1]; u[2 * r] = a + b; u[2 * r + 1] = a - b; }
  }
}
// forward-real epilogue on PAIRS: X[k] and X[Nc-k] share s = Z[k] + conj Z[Nc-k] and u = W^k (Z[k] - conj Z[Nc-k])
//   X[k] = ((s.x + u.y), (s.y - u.x))/2      X[Nc-k] = ((s.x - u.y), (-s.y - u.x))/2
// k = 0 also emits the self-paired middle bin: slot 0 = (Z0.x + Z0.y, Z0.x - Z0.y), X[Nc/2] = conj Z[Nc/2].
template <int SM, typename T, bool SWZ = false>
PF_HD void real_post_pair(T* base, const cpx<T>* z, int k, int N, int Nc, const cpx<T>* twr) {
  constexpr bool Z = (SM == S_R_Z);
  if (k == 0) {
    const cpx<T> z0 = z[0], zm = z[Nc / 2];
    spec_put<Z, true, SWZ>(base, 0, N, mk<T>(z0.x + z0.y, z0.x - z0.y));
    spec_put<Z, true, SWZ>(base, Nc / 2, N, mk<T>(zm.x, -zm.y));
    return;
  }
  const cpx<T> a = z[k], b = conj(z[Nc - k]);
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul(d, ldtab(twr + k));
  spec_put<Z, true, SWZ>(base, k, N, scale(s + mul_mi(u), T(0.5)));
  spec_put<Z, true, SWZ>(base, Nc - k, N, emul(s - mul_mi(u), mk<T>(T(0.5), T(-0.5))));
}


// ---- forward-real variant of pass 3 for C <= 8: each thread takes, besides its row (k_a, k_b), the row that holds
// the mirror bins Nc-k, so the real post-rotation happens in registers (no natural-order round trip through shared
// memory, no extra barriers).  Mirror of k = k_a + 16 k_b + 256 k_c:
//     k_a != 0          : (16-k_a, 15-k_b, C-1-k_c)
//     k_a == 0, k_b != 0: (0, 16-k_b, C-1-k_c)
//     k_a == 0, k_b == 0: inside row (0,0) [k_c <-> C-k_c]; thread 0 pairs it with row (0,8), whose mirror is itself
// Representative rows are those with k_b < 8: thread t -> k_a = t%16, k_b = t/16 + C*r, r < 8/C.
template <int C> PF_HD void k2_mirror_row(int ka, int kb, int& ka2, int& kb2) {
  if (ka != 0) { ka2 = 16 - ka; kb2 = 15 - kb; }
  else { ka2 = 0; kb2 = (kb == 0) ? 8 : 16 - kb; }
}
template <int C, int SIGN, typename T>
PF_HD void k2_pass3_pairs(int t, const cpx<T>* tile, cpx<T> (&u)[16]) {
  using K = K2<C>;
  const int ka = t & 15, kb0 = t >> 4;
#pragma unroll
  for (int r = 0; r < 8 / C; ++r) {
    const int kb = kb0 + C * r;
    int ka2, kb2;
    k2_mirror_row<C>(ka, kb, ka2, kb2);
#pragma unroll
    for (int p = 0; p < C; ++p) {
      u[(2 * r) * C + p] = tile[K::idx(ka, kb, brevC<C>(p))];
      u[(2 * r + 1) * C + p] = tile[K::idx(ka2, kb2, brevC<C>(p))];
    }
  }
  if (C == 8) { dit_fft<8, SIGN, 0, 1>(u); dit_fft<8, SIGN, 8, 1>(u); }
  if (C == 4) { dit_fft<4, SIGN, 0, 1>(u); dit_fft<4, SIGN, 4, 1>(u); dit_fft<4, SIGN, 8, 1>(u); dit_fft<4, SIGN, 12, 1>(u); }
  if (C == 2) {
#pragma unroll
    for (int r = 0; r < 8; ++r) { const cpx<T> a = u[2 * r], b = u[2 * r + 1]; u[2 * r] = a + b; u[2 * r + 1] = a - b; }
  }
}
// X[k], X[Nc-k] from a = Z[k], zm = Z[Nc-k]  (same algebra as real_post_pair)
template <int SM, typename T, bool SWZ = false>
PF_HD void real_post_regs(T* base, int k, int Nc, int N, cpx<T> a, cpx<T> zm, const cpx<T>* twr) {
  constexpr bool Z = (SM == S_R_Z);
  const cpx<T> b = conj(zm);
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul(d, ldtab(twr + k));
  spec_put<Z, true, SWZ>(base, k, N, scale(s + mul_mi(u), T(0.5)));
  spec_put<Z, true, SWZ>(base, Nc - k, N, emul(s - mul_mi(u), mk<T>(T(0.5), T(-0.5))));
}
template <int C, int SM, typename T, bool SWZ = false>
PF_HD void k2_store_pairs(int t, const cpx<T> (&u)[16], T* base, int N, const cpx<T>* twr) {
  using K = K2<C>;
  constexpr bool Z = (SM == S_R_Z);
  const int ka = t & 15, kb0 = t >> 4;
#pragma unroll
  for (int r = 0; r < 8 / C; ++r) {
    const int kb = kb0 + C * r;
    const cpx<T>* A = &u[(2 * r) * C];
    const cpx<T>* B = &u[(2 * r + 1) * C];
    if (ka == 0 && kb == 0) {
      // row (0,0): k = 256 kc, mirror inside the row; row (0,8): k = 128 + 256 kc, mirror inside that row
      spec_put<Z, true, SWZ>(base, 0, N, mk<T>(A[0].x + A[0].y, A[0].x - A[0].y));       // (DC, Nyquist)
      if (C >= 2) spec_put<Z, true, SWZ>(base, K::NC / 2, N, mk<T>(A[C / 2].x, -A[C / 2].y));  // X[Nc/2] = conj Z[Nc/2]
#pragma unroll
      for (int kc = 1; kc < C / 2; ++kc) real_post_regs<SM, T, SWZ>(base, 256 * kc, K::NC, N, A[kc], A[C - kc], twr);
#pragma unroll
      for (int kc = 0; kc < C / 2; ++kc) real_post_regs<SM, T, SWZ>(base, 128 + 256 * kc, K::NC, N, B[kc], B[C - 1 - kc], twr);
    } else {
#pragma unroll
      for (int kc = 0; kc < C; ++kc) real_post_regs<SM, T, SWZ>(base, ka + 16 * kb + 256 * kc, K::NC, N, A[kc], B[C - 1 - kc], twr);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BACKWARD REAL (ordered spectrum in, time samples out) for C <= 8 as the MIRROR IMAGE of the forward passes -- decimation
// in frequency (round 2b).  The element-wise form (k2_pass1<L_R_ORD>) rebuilds every packed input Z[n] from X[n] and X[Nc-n]
// with two mirrored global reads and a rotation per element: 2x the load instructions and LSU wavefronts of the forward
// kernel (C3 backward 0.75 of the roofline against 0.83 forward).  Mirrored, the thread that owns row (k_a, k_b) also owns
// the row with the mirror bins (the forward kernel's k2_store_pairs, read backwards): every spectrum element is loaded ONCE
// and a pair (k, Nc-k) shares its sum, difference and twiddle:
//     a = X[k], b = conj X[Nc-k], s = a + b, u = (a - b) * conj(w_k):   Z[k] = s + i u,   Z[Nc-k] = conj(s) + i conj(u)
//   pass B1: thread (k_a, k_b) [+ mirror row]: pre-rotation in registers, radix-C DFT over k_c -> n_c, * conj W_{16C}^{n_c k_b}
//   pass B2: thread (k_a, n_c): radix 16 over k_b -> n_b, * conj W_Nc^{(n_c + C n_b) k_a}            (in place)
//   pass B3: thread m = (n_b, n_c): radix 16 over k_a -> n_a, stores x[m + 16C n_a]                  (coalesced)
// Same tile, same swizzle: B1 writes what forward pass 3 reads, B3 reads what forward pass 1 writes.
// ---------------------------------------------------------------------------------------------------------------
// Z[k], Z[Nc-k] from a = X[k], xm = X[Nc-k]
template <typename T> PF_HD void real_pre_regs(cpx<T> a, cpx<T> xm, cpx<T> w, cpx<T>* zk, cpx<T>* zm) {
  const cpx<T> b = conj(xm);
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul_dir<+1>(d, w);                       // d * conj(w)
  *zk = s + mul_pi(u);                                        // s + i u
  *zm = conj(s) + mul_pi(conj(u));                            // conj(s) + i conj(u)
}
template <int C, typename T>
PF_HD void k2b_pass1_pairs(int t, const T* base, int N, const cpx<T>* twr, const cpx<T>* tw2, cpx<T>* tile) {
  using K = K2<C>;
  const cpx<T>* X = reinterpret_cast<const cpx<T>*>(base);
  const int ka = t & 15, kb0 = t >> 4;
#pragma unroll
  for (int r = 0; r < 8 / C; ++r) {
    const int kb = kb0 + C * r;
    int ka2, kb2;
    k2_mirror_row<C>(ka, kb, ka2, kb2);
    cpx<T> xa[C], xb[C], za[C], zb[C];
#pragma unroll
    for (int kc = 0; kc < C; ++kc) { xa[kc] = X[ka + 16 * kb + 256 * kc]; xb[kc] = X[ka2 + 16 * kb2 + 256 * kc]; }
    if (ka == 0 && kb == 0) {
      // row (0,0): k = 256 kc pairs with 256 (C - kc) inside the row; row (0,8): k = 128 + 256 kc with 128 + 256 (C-1-kc)
      za[0] = mk<T>(xa[0].x + xa[0].y, xa[0].x - xa[0].y);                          // slot 0 = (DC, Nyquist)
      if (C >= 2) za[C / 2] = scale(conj(xa[C / 2]), T(2));                          // Z[Nc/2] = 2 conj X[Nc/2]
#pragma unroll
      for (int kc = 1; kc < C / 2; ++kc) real_pre_regs<T>(xa[kc], xa[C - kc], ldtab(twr + 256 * kc), &za[kc], &za[C - kc]);
#pragma unroll
      for (int kc = 0; kc < C / 2; ++kc) real_pre_regs<T>(xb[kc], xb[C - 1 - kc], ldtab(twr + 128 + 256 * kc), &zb[kc], &zb[C - 1 - kc]);
    } else {
#pragma unroll
      for (int kc = 0; kc < C; ++kc) real_pre_regs<T>(xa[kc], xb[C - 1 - kc], ldtab(twr + ka + 16 * kb + 256 * kc), &za[kc], &zb[C - 1 - kc]);
    }
    dft_small<C, +1>(za);
    dft_small<C, +1>(zb);
    tile[K::idx(ka, kb, 0)] = za[0];
    tile[K::idx(ka2, kb2, 0)] = zb[0];
#pragma unroll
    for (int nc = 1; nc < C; ++nc) {
      tile[K::idx(ka, kb, nc)] = cmul_dir<+1>(za[nc], ldtab(tw2 + kb * C + nc));
      tile[K::idx(ka2, kb2, nc)] = cmul_dir<+1>(zb[nc], ldtab(tw2 + kb2 * C + nc));
    }
  }
}
template <int C, typename T>
PF_HD void k2b_pass2(int t, const cpx<T>* tw1, cpx<T>* tile) {
  using K = K2<C>;
  const int ka = t / C, nc = t % C;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = tile[K::idx(ka, brev4(p), nc)];
  reg_fft<16, +1>(v);
  if (ka == 0) {
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) tile[K::idx(ka, nb, nc)] = v[nb];
  } else {
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) tile[K::idx(ka, nb, nc)] = cmul_dir<+1>(v[nb], ldtab(tw1 + ka * K::BC + nb * C + nc));
  }
}
template <int C, typename T>
PF_HD void k2b_pass3(int m, const cpx<T>* tile, cpx<T> (&v)[16]) {
  using K = K2<C>;
  const int jb = m / C, jc = m % C;
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = tile[K::idx(brev4(p), jb, jc)];
  reg_fft<16, +1>(v);                                         // v[n_a] = x[m + 16C n_a] (complex pair of real samples)
}

// natural index of u[r*C + kc] held by thread t after pass 3
template <int C> PF_HD int k2_out_index(int t, int r, int kc) { return (t & 15) + 16 * ((t >> 4) + C * r) + 256 * kc; }

#ifdef __CUDACC__
// One CTA = one transform at a time, persistent over the batch.  blockDim.x == 16*C.
// STAGED: the input of the NEXT transform is fetched by the TMA engine (1-D cp.async.bulk, SASS UBLKCP) into a
// second shared buffer while passes 2/3 and the stores of the current one run; pass 1 then reads shared memory.
// Only for contiguous, 16-byte aligned, fully in-range inputs in canonical order (complex or real time samples).
// 16-byte granule copies between a dense global spectrum and its shared staging copy (optionally z-swizzled)
template <typename T, int NTHR, bool SWZ> PF_D void stage_in16(const T* g, T* sm, int nelem, int t) {
  constexpr int EPG = 16 / sizeof(T);                      // elements per 16-byte granule
  for (int G = t; G < nelem / EPG; G += NTHR) {
    const int e = G * EPG;
    *reinterpret_cast<float4*>(sm + (SWZ ? zswz(e) : e)) = *reinterpret_cast<const float4*>(g + e);
  }
}
template <typename T, int NTHR, bool SWZ> PF_D void stage_out16(T* g, const T* sm, int nelem, int t) {
  constexpr int EPG = 16 / sizeof(T);
  for (int G = t; G < nelem / EPG; G += NTHR) {
    const int e = G * EPG;
    *reinterpret_cast<float4*>(g + e) = *reinterpret_cast<const float4*>(sm + (SWZ ? zswz(e) : e));
  }
}
PF_HD bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// One CTA = one transform at a time, persistent over the batch.  blockDim.x == 16*C.
// STAGED: the input of the NEXT transform is fetched by the TMA engine (1-D cp.async.bulk, SASS UBLKCP) into a
// second shared buffer while passes 2/3 and the stores of the current one run; pass 1 then reads shared memory.
// Only for contiguous, 16-byte aligned, fully in-range inputs in canonical order (complex or real time samples).
// Inputs whose elements are gathered (z-domain layouts, backward-real pre-rotation: X[k] and X[Nc-k]) are first
// copied with coalesced 128-bit loads into a second shared buffer, and z-domain outputs are assembled in shared
// memory and written with coalesced 128-bit stores, both with the granule swizzle `zswz`.
template <typename T, int C, int LM, int SM, int SIGN, int MINB, bool STAGED>
__global__ void __launch_bounds__(16 * C, MINB)
k_cta_fft(const XformParams<T> p, const cpx<T>* tw1, const cpx<T>* tw2) {
  using K = K2<C>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* stage = tile + K::NC;                                   // STAGED or kGatherIn
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + K::NC);     // STAGED only
  const int t = threadIdx.x;
  constexpr bool kNeedsPartner = (SM == S_R_ORD || SM == S_R_Z);  // forward real: X[k] needs Z[k] and Z[Nc-k]
  // (staging the ORDERED backward-real input was measured slower than its direct mirrored loads: 0.58 vs 0.75)
  constexpr bool kGatherIn = (LM == L_C_Z || LM == L_R_Z);
  constexpr bool kMirrorBwd = (LM == L_R_ORD && SM == S_R_TIME && SIGN > 0 && C <= 8 && !STAGED);
  // the API length is fixed by the kernel: a compile-time N turns the z-domain index maps into shifts and masks
  constexpr int kN = (LM == L_C_ORD || LM == L_C_Z) ? K::NC : 2 * K::NC;
  constexpr bool kZIn = (LM == L_C_Z || LM == L_R_Z);
  constexpr bool kZOut = (SM == S_C_Z || SM == S_R_Z);
  constexpr uint32_t kStageBytes = K::NC * sizeof(cpx<T>);
  if (STAGED) {
    if (t == 0) {
      mbar_init(bar, 1); fence_mbar_init(); fence_proxy_async();
      if ((long long)blockIdx.x < p.batch) { mbar_expect_tx(bar, kStageBytes); bulk_g2s(stage, p.in + (long long)blockIdx.x * p.in_stride, kStageBytes, bar); }
    }
    __syncthreads();
  }
  uint32_t phase = 0;
  for (long long tr = blockIdx.x; tr < p.batch; tr += gridDim.x) {
    // keep the twiddle loads inside the loop: hoisted, the 30 per-thread twiddles are loop invariants the
    // compiler spills to local memory (measured: +2 GB of L2 traffic per 4 GB launch); re-read from L1 instead
    const cpx<T>* twr = p.twr;
    asm volatile("" : "+l"(tw1), "+l"(tw2), "+l"(twr));
    const long long grp = (p.in_group > 1) ? tr / p.in_group : tr;
    const T* ibase = p.in + grp * p.in_stride + (tr - grp * p.in_group) * (long long)p.in_gstep;
    T* obase = p.out + tr * p.out_stride;
    if (STAGED) {
      mbar_wait(bar, phase); phase ^= 1;
      k2_pass1<C, LM, SIGN, true, T>(t, reinterpret_cast<const T*>(stage), kN, twr, -1, true, tw1, tile);
      __syncthreads();
      const long long nxt = tr + gridDim.x;
      if (t == 0 && nxt < p.batch) {                                   // stage is free: fetch the next transform now
        fence_proxy_async();
        mbar_expect_tx(bar, kStageBytes);
        bulk_g2s(stage, p.in + nxt * p.in_stride, kStageBytes, bar);
      }
    } else if (kMirrorBwd && p.in_estride == 1 && p.in_limit < 0 && vec_aligned<T>(ibase) && vec_aligned<T>(obase)) {
      // backward real, ordered spectrum: decimation in frequency with in-register pair pre-rotation (see k2b_pass1_pairs)
      k2b_pass1_pairs<C, T>(t, ibase, kN, twr, tw2, tile);
      __syncthreads();
      k2b_pass2<C, T>(t, tw1, tile);
      __syncthreads();
      cpx<T> v[16];
      k2b_pass3<C, T>(t, tile, v);
      if (p.out_count >= 2 * K::NC) {
        cpx<T>* dst = reinterpret_cast<cpx<T>*>(obase);
#pragma unroll
        for (int na = 0; na < 16; ++na) dst[t + K::BC * na] = v[na];
      } else {
#pragma unroll
        for (int na = 0; na < 16; ++na) store_elem<SM, T>(obase, t + K::BC * na, v[na], kN, p.out_count, true);
      }
      __syncthreads();                          // tile is rewritten by the next transform
      continue;
    } else if (kGatherIn && aligned16(ibase)) {
      // whole spectrum -> shared (coalesced), then every thread gathers its 16 points (and their mirrors) from there
      stage_in16<T, 16 * C, kZIn>(ibase, reinterpret_cast<T*>(stage), 2 * K::NC, t);
      __syncthreads();
      k2_pass1<C, LM, SIGN, false, T, kZIn>(t, reinterpret_cast<const T*>(stage), kN, twr, -1, true, tw1, tile);
      __syncthreads();
    } else {
      const long long avail = (p.in_limit < 0) ? -1 : (p.in_limit - tr * p.in_stride);
      const bool vin = vec_aligned<T>(ibase);
      if (vin && p.in_estride == 1 && (avail < 0 || avail >= (long long)(2 * K::NC)))
        k2_pass1<C, LM, SIGN, true, T>(t, ibase, kN, twr, avail, true, tw1, tile);
      else
        k2_pass1<C, LM, SIGN, false, T>(t, ibase, kN, twr, avail, vin, tw1, tile, p.in_estride);
      __syncthreads();
    }
    k2_pass2<C, SIGN, T>(t, tw2, tile);
    __syncthreads();
    cpx<T> u[16];
    const bool zstage = kZOut && aligned16(obase) && !(STAGED && C == 16 && SM == S_R_Z);   // CTA-uniform
    if (kNeedsPartner && C <= 8) {             // forward real, mirror rows in the same thread: rotate in registers
      k2_pass3_pairs<C, SIGN, T>(t, tile, u);
      if (zstage) {
        __syncthreads();                        // all pass-3 reads done: the tile becomes the z-domain staging buffer
        k2_store_pairs<C, SM, T, true>(t, u, reinterpret_cast<T*>(tile), kN, twr);
        __syncthreads();
        stage_out16<T, 16 * C, true>(obase, reinterpret_cast<const T*>(tile), 2 * K::NC, t);
      } else {
        k2_store_pairs<C, SM, T>(t, u, obase, kN, twr);
      }
      __syncthreads();
      continue;
    }
    k2_pass3<C, SIGN, T>(t, tile, u);
    if (!kNeedsPartner) {
      const bool vok = vec_aligned<T>(obase);
      if (SM == S_C_ORD || (SM == S_R_TIME && vok && p.out_count >= 2 * K::NC)) {   // whole transform stored: no per-element checks
        cpx<T>* dst = reinterpret_cast<cpx<T>*>(obase);
#pragma unroll
        for (int r = 0; r < 16 / C; ++r)
#pragma unroll
          for (int kc = 0; kc < C; ++kc) dst[k2_out_index<C>(t, r, kc)] = u[r * C + kc];
      } else if (SM == S_C_Z && zstage) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16 / C; ++r)
#pragma unroll
          for (int kc = 0; kc < C; ++kc) store_elem<SM, T, true>(reinterpret_cast<T*>(tile), k2_out_index<C>(t, r, kc), u[r * C + kc], kN, p.out_count, true);
        __syncthreads();
        stage_out16<T, 16 * C, true>(obase, reinterpret_cast<const T*>(tile), 2 * K::NC, t);
      } else {
#pragma unroll
        for (int r = 0; r < 16 / C; ++r)
#pragma unroll
          for (int kc = 0; kc < C; ++kc) store_elem<SM, T>(obase, k2_out_index<C>(t, r, kc), u[r * C + kc], kN, p.out_count, vok);
      }
      __syncthreads();                          // tile is rewritten by the next transform's pass 1
    } else {
      __syncthreads();                          // everyone has read the tile: reuse it as the natural-order buffer
#pragma unroll
      for (int r = 0; r < 16 / C; ++r)
#pragma unroll
        for (int kc = 0; kc < C; ++kc) tile[k2_out_index<C>(t, r, kc)] = u[r * C + kc];
      __syncthreads();
      if (zstage) {
        // C == 16: natural-order spectrum in `tile`; assemble the z-domain image in `stage`, then stream it out
#pragma unroll 4
        for (int j = 0; j < 8; ++j) real_post_pair<SM, T, true>(reinterpret_cast<T*>(stage), tile, t + K::T * j, kN, K::NC, twr);
        __syncthreads();
        stage_out16<T, 16 * C, true>(obase, reinterpret_cast<const T*>(stage), 2 * K::NC, t);
      } else {
#pragma unroll 4
        for (int j = 0; j < 8; ++j) real_post_pair<SM, T>(obase, tile, t + K::T * j, kN, K::NC, twr);   // k in [0, Nc/2)
      }
      __syncthreads();
    }
  }
}
#endif  // __CUDACC__


// ---------------------------------------------------------------------------------------------------------------
// Two-level plans that still fit one CTA: Nc = R x N2 with N2 = 256*C.  The CTA runs the R row transforms
// (decimated sub-sequences x[n1 + R*n2], the three register passes above) one after the other, parks each row,
// already multiplied by W_Nc^{n1 k}, in shared memory, and finishes with radix-R register DFTs across the rows:
//     X[k2 + N2*k1] = sum_n1 W_R^{n1 k1} * ( W_Nc^{n1 k2} * Y_n1[k2] )
// One HBM read and one HBM write per transform instead of the two round trips of the split (rows + combine) plan.
// twP[n1*N2 + k] = exp(-2 pi i n1 k / Nc) is the row-major copy of the combine twiddles (unit-stride reads per row).
// Shared memory: (R + 1) * N2 complex words.  Complex canonical in and out only (other layouts wrap it).
// ---------------------------------------------------------------------------------------------------------------
// park one finished row: thread t holds u[r*C + kc] = Y_n1[k] (k = k2_out_index); slot of (n1, k) is rows[n1*RS + k*KS]
// (row-major: RS = N2, KS = 1)
template <int C, int R, int SIGN, int RS, int KS, typename T>
PF_HD void split_park(int t, int n1, const cpx<T> (&u)[16], const cpx<T>* twP, cpx<T>* rows) {
  constexpr int N2 = K2<C>::NC;
#pragma unroll
  for (int r = 0; r < 16 / C; ++r)
#pragma unroll
    for (int kc = 0; kc < C; ++kc) {
      const int k = k2_out_index<C>(t, r, kc);
      rows[n1 * RS + k * KS] = (n1 == 0) ? u[r * C + kc] : cmul_dir<SIGN>(u[r * C + kc], ldtab(twP + n1 * N2 + k));
    }
}
template <int C, int R, int SIGN, int RS, int KS, typename T>
PF_HD void split_combine_cols(int t, const cpx<T>* rows, cpx<T>* dst) {
  using K = K2<C>;
  constexpr int N2 = K::NC;
#pragma unroll 2
  for (int j = 0; j < 16; ++j) {
    const int k2 = t + K::T * j;
    cpx<T> v[R];
#pragma unroll
    for (int n1 = 0; n1 < R; ++n1) v[n1] = rows[n1 * RS + k2 * KS];
    dft_small<R, SIGN>(v);
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) dst[k2 + N2 * k1] = v[k1];
  }
}

#ifdef __CUDACC__
// (Measured and dropped: copying the whole transform into shared memory with dense 128-bit loads first and parking the
//  rows in place over that copy -- HBM/L2 then see no strided reads -- was SLOWER for every size but one: 8192: 0.47 vs
//  0.57, 3072: 0.49 vs 0.73, 9216: 0.32 vs 0.43 of HBM peak.  L2 absorbs the strided row reads; the extra shared-memory
//  round trip and barrier do not pay.)
template <typename T, int C, int R, int SIGN, int MINB>
__global__ void __launch_bounds__(16 * C, MINB)
k_cta_split(const T* in, T* out, long long batch, const cpx<T>* tw1, const cpx<T>* tw2, const cpx<T>* twP) {
  using K = K2<C>;
  constexpr int N2 = K::NC;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* rows = tile + N2;                                   // [R][N2]
  const int t = threadIdx.x;
  for (long long tr = blockIdx.x; tr < batch; tr += gridDim.x) {
    asm volatile("" : "+l"(tw1), "+l"(tw2), "+l"(twP));
    const cpx<T>* src = reinterpret_cast<const cpx<T>*>(in) + tr * (long long)(R * N2);
    cpx<T>* dst = reinterpret_cast<cpx<T>*>(out) + tr * (long long)(R * N2);
#pragma unroll 1
    for (int n1 = 0; n1 < R; ++n1) {
      k2_pass1<C, L_C_ORD, SIGN, false, T>(t, reinterpret_cast<const T*>(src + n1), N2, nullptr, -1, true, tw1, tile, R);
      __syncthreads();
      k2_pass2<C, SIGN, T>(t, tw2, tile);
      __syncthreads();
      cpx<T> u[16];
      k2_pass3<C, SIGN, T>(t, tile, u);
      split_park<C, R, SIGN, N2, 1, T>(t, n1, u, twP, rows);
      __syncthreads();                                        // tile free for the next row; row n1 complete
    }
    split_combine_cols<C, R, SIGN, N2, 1, T>(t, rows, dst);
    __syncthreads();                                          // rows are rewritten by the next transform
  }
}
#endif  // __CUDACC__

}  // namespace pf
