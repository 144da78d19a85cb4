// radix_kernels.cuh -- ONE transform (or a group of small ones) per CTA for complex cores that are neither 32*R2 (warp
// kernels) nor 256*C (16x16xC kernels): 16, 48, 80, 144, 240, 400, 432, 1296, 2000, 2592, 4000, 6000, 12000 ... -- the sizes the
// reference's validator walks (benchmarks/bench_pffft.c:445) that were left to the runtime-radix shared-memory kernel.
//
// Nc = R1 * R2 [* R3], every radix a register DFT of the library (butterfly.cuh), all COMPILE-TIME.  Autosort (Stockham)
// stages with ONE butterfly per thread and stage:
//   stage 1: thread b < Nc/R1 reads x[b + j*Nc/R1] STRAIGHT FROM GLOBAL memory (coalesced over b; the load function also
//            performs the backward-real pre-rotation and the z-domain gather), radix-R1 DFT, * W_Nc^{b k}, to shared memory;
//   stage 2: shared -> registers -> radix-R2 -> * W_Nc^{R1 p k} -> shared (in place, one barrier between reads and writes);
//   stage 3: shared -> registers -> radix-R3 -> STRAIGHT TO GLOBAL memory (coalesced over b).  Forward-real outputs, which
//            need the mirror bin, take one more trip through shared memory.
// One HBM read and one HBM write per transform, two (three) barriers, no runtime radix dispatch, no index division that is
// not by a constant.  The stage-1 result is stored with one pad word per R1 words when R1 is even, which makes the
// stride-R1 writes of neighbouring threads conflict free (odd stride in 8-byte words).
//
// Replaces, for these sizes, passf2/3/4/5_ps + cplx_finalize (complex) and radf*/radb* + real_finalize/preprocess (real)
// of the reference (src/pffft_priv_impl.h:122-807, :1195-1462).
#pragma once
#ifndef RADIX_NO_PAIRS
#define RADIX_NO_PAIRS 0                                          // 1: forward real through the natural-order copy + pairing pass (round 2)
#endif
#include "butterfly.cuh"
#include "generic_kernels.cuh"
#include "cta_kernels.cuh"   // store_elem, ldtab

namespace pf {

template <int R1, int R2, int R3> struct RadixShape {
  static constexpr int NC = R1 * R2 * R3;
  static constexpr int M1 = NC / R1, M2 = NC / R2, M3 = NC / R3;
  static constexpr int STAGES = R3 > 1 ? 3 : (R2 > 1 ? 2 : 1);
  static constexpr int TT = (STAGES == 3) ? (M3 > M2 ? (M3 > M1 ? M3 : M1) : (M2 > M1 ? M2 : M1))
                          : (STAGES == 2) ? (M2 > M1 ? M2 : M1) : M1;          // threads per transform
  static constexpr int PAD = (R1 % 2 == 0) ? 1 : 0;
  static constexpr int NCP = NC + PAD * (NC / R1) + 8;                          // words of one transform's buffer
  PF_HD static int idx1(int n) { return PAD ? n + n / R1 : n; }                // layout of the stage-1 result
};

// W_Nc^{e} for SIGN (table holds the forward roots)
template <int SIGN, typename T> PF_HD cpx<T> radix_tw(cpx<T> v, const cpx<T>* tw, int e) { return cmul_dir<SIGN>(v, ldtab(tw + e)); }

// ---- backward real: the packed half-length spectrum Z' = 2 (E + i O) is built PAIR-WISE in shared memory first --
// bins k and Nc-k share their sum, difference and twiddle -- instead of letting every stage-1 load rebuild
// its element from two mirrored global reads.  `nat` receives Z' in natural order.
template <typename T, int LM>
PF_HD void radix_prerotate(int li, int TT, const T* ibase, int N, int Nc, const cpx<T>* twr, cpx<T>* nat) {
  constexpr bool Z = (LM == L_R_Z);
  for (int k = li; k < Nc / 2; k += TT) {
    if (k == 0) {
      const cpx<T> s0 = spec_get<Z, true>(ibase, 0, N);                 // (X[0], X[N/2])
      nat[0] = mk<T>(s0.x + s0.y, s0.x - s0.y);
      const cpx<T> am = spec_get<Z, true>(ibase, Nc / 2, N);            // self-paired middle bin: Z'[Nc/2] = 2 conj X[Nc/2]
      nat[Nc / 2] = mk<T>(T(2) * am.x, T(-2) * am.y);
      continue;
    }
    const cpx<T> a = spec_get<Z, true>(ibase, k, N);
    const cpx<T> b = conj(spec_get<Z, true>(ibase, Nc - k, N));
    const cpx<T> s = a + b, d = a - b;
    const cpx<T> u = cmul_dir<+1>(d, ldtab(twr + k));                   // d * exp(+2 pi i k / N)
    nat[k] = mk<T>(s.x - u.y, s.y + u.x);                               // s + i u
    nat[Nc - k] = mk<T>(s.x + u.y, u.x - s.y);                          // conj(s) + i conj(u)
  }
}

// ---- the three stages of ONE transform for thread li (0 <= li < TT); `buf` = this transform's shared buffer
// stage 1, first half: the R1 inputs of butterfly li -- from global memory (any load mode) or from the natural-order
// shared buffer a pre-rotation left behind
template <typename T, int R1, int R2, int R3, int LM, bool FROM_SMEM>
PF_HD void radix_stage1_load(int li, const T* ibase, int N, const cpx<T>* twr, const cpx<T>* nat, cpx<T> (&a)[R1]) {
  using S = RadixShape<R1, R2, R3>;
  if (li >= S::M1) return;
#pragma unroll
  for (int j = 0; j < R1; ++j) a[j] = FROM_SMEM ? nat[li + j * S::M1] : load_core<LM, T>(ibase, li + j * S::M1, N, S::NC, twr, -1, true);
}
// second half: DFT, * W_Nc^{b k}, into the padded stage-1 layout (single-stage plans keep the result in `a`)
template <typename T, int R1, int R2, int R3, int SIGN>
PF_HD void radix_stage1_store(int li, cpx<T> (&a)[R1], const cpx<T>* tw, cpx<T>* buf) {
  using S = RadixShape<R1, R2, R3>;
  if (li >= S::M1) return;
  dft_small<R1, SIGN>(a);
  if (S::STAGES == 1) return;
  buf[S::idx1(R1 * li)] = a[0];
#pragma unroll
  for (int k = 1; k < R1; ++k) buf[S::idx1(R1 * li + k)] = radix_tw<SIGN>(a[k], tw, li * k);
}
template <typename T, int R1, int R2, int R3>
PF_HD void radix_stage2_read(int li, const cpx<T>* buf, cpx<T> (&a)[R2]) {
  using S = RadixShape<R1, R2, R3>;
  if (li >= S::M2) return;
#pragma unroll
  for (int j = 0; j < R2; ++j) a[j] = buf[S::idx1(li + j * S::M2)];
}
// stage 2 of a 3-stage plan: in place (caller put a barrier between read and write)
template <typename T, int R1, int R2, int R3, int SIGN>
PF_HD void radix_stage2_write(int li, cpx<T> (&a)[R2], const cpx<T>* tw, cpx<T>* buf) {
  using S = RadixShape<R1, R2, R3>;
  if (li >= S::M2) return;
  dft_small<R2, SIGN>(a);
  const int p = li / R1, q = li - p * R1;
  cpx<T>* o = buf + q + R1 * R2 * p;
  o[0] = a[0];
#pragma unroll
  for (int k = 1; k < R2; ++k) o[R1 * k] = p ? radix_tw<SIGN>(a[k], tw, R1 * p * k) : a[k];
}
template <typename T, int R1, int R2, int R3>
PF_HD void radix_stage3_read(int li, const cpx<T>* buf, cpx<T> (&a)[R3]) {
  using S = RadixShape<R1, R2, R3>;
  if (li >= S::M3) return;
#pragma unroll
  for (int j = 0; j < R3; ++j) a[j] = buf[li + j * S::M3];
}

// emit the LAST stage's outputs v[k] = X[b + M*k] (M = Nc / R_last, thread b < M): straight to global memory for the modes
// whose element needs no partner, else into `nat` (natural order, shared) for the pairing pass
template <typename T, int RL, int M, int SM>
PF_HD void radix_emit(int b, const cpx<T> (&v)[RL], T* obase, int N, cpx<T>* nat) {
  constexpr bool partner = (SM == S_R_ORD || SM == S_R_Z);
  if (b >= M) return;
#pragma unroll
  for (int k = 0; k < RL; ++k) {
    if (partner) nat[b + M * k] = v[k];
    else store_elem<SM, T>(obase, b + M * k, v[k], N, N, true);
  }
}

// ---- forward real, last stage on PAIRS of butterflies (round 2b).  Thread p <= M/2 runs the butterflies b = p and M - p: the
// outputs of the second one, X[(M-p) + M k], are exactly the mirror bins Nc - (p + M k') of the first (k' = R-1-k), so the real
// post-rotation happens in registers and both halves of every pair go straight to global memory -- no natural-order copy in
// shared memory, no pairing pass, two barriers less.  p = 0 and p = M/2 pair inside their own butterfly.  Used when the last
// radix is <= 16 (two butterflies = 4 R registers of data).  PADDED: the stage-1 layout (two-stage plans read it).
template <typename T, int R, int M, int NC, int SM, bool PADDED, int R1>
PF_HD void radix_last_pairs(int p, const cpx<T>* buf, T* obase, int N, const cpx<T>* twr) {
  if (2 * p > M) return;
  const int b2 = (p == 0) ? 0 : M - p;
  auto rd = [&](int n) { return PADDED ? buf[n + n / R1] : buf[n]; };
  cpx<T> v1[R], v2[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { v1[j] = rd(p + j * M); v2[j] = rd(b2 + j * M); }
  dft_small<R, -1>(v1);
  if (p == 0) {                                                  // n = M k  <->  M (R - k)
    spec_put<SM == S_R_Z, true>(obase, 0, N, mk<T>(v1[0].x + v1[0].y, v1[0].x - v1[0].y));          // (DC, Nyquist)
    if (R % 2 == 0) spec_put<SM == S_R_Z, true>(obase, NC / 2, N, mk<T>(v1[R / 2].x, -v1[R / 2].y));   // X[Nc/2] = conj Z[Nc/2]
#pragma unroll
    for (int k = 1; 2 * k < R; ++k) real_post_regs<SM, T>(obase, M * k, NC, N, v1[k], v1[R - k], twr);
    return;
  }
  if (2 * p == M) {                                              // n = M/2 + M k  <->  M/2 + M (R-1-k)
#pragma unroll
    for (int k = 0; 2 * k < R - 1; ++k) real_post_regs<SM, T>(obase, p + M * k, NC, N, v1[k], v1[R - 1 - k], twr);
    if (R % 2 == 1) { const int k = (R - 1) / 2; spec_put<SM == S_R_Z, true>(obase, p + M * k, N, mk<T>(v1[k].x, -v1[k].y)); }   // self-paired: n = Nc/2
    return;
  }
  dft_small<R, -1>(v2);
#pragma unroll
  for (int k = 0; k < R; ++k) real_post_regs<SM, T>(obase, p + M * k, NC, N, v1[k], v2[R - 1 - k], twr);
}

// ---- backward real, FIRST stage on pairs of butterflies: the mirror image of radix_last_pairs.  Thread p <= M1/2 loads the inputs
// of the butterflies b = p and M1 - p straight from the spectrum: element j of the second is the mirror bin of element
// R1-1-j of the first, so the pre-rotation (real_pre_regs, cta_kernels.cuh) runs in registers -- no natural-order copy of the
// rebuilt half-length spectrum in shared memory, two barriers less.  Then both butterflies take the ordinary stage-1 path.
template <typename T, int R1, int R2, int R3, int LM>
PF_HD void radix_first_pairs(int p, const T* ibase, int N, const cpx<T>* twr, const cpx<T>* tw, cpx<T>* buf) {
  using S = RadixShape<R1, R2, R3>;
  constexpr int M = S::M1, NC = S::NC;
  constexpr bool Z = (LM == L_R_Z);
  if (2 * p > M) return;
  const int b2 = (p == 0) ? 0 : M - p;
  cpx<T> x1[R1], z1[R1];
#pragma unroll
  for (int j = 0; j < R1; ++j) x1[j] = spec_get<Z, true>(ibase, p + j * M, N);
  if (p == 0) {                                                  // k = M j  <->  M (R1 - j)
    z1[0] = mk<T>(x1[0].x + x1[0].y, x1[0].x - x1[0].y);        // slot 0 = (DC, Nyquist)
    if (R1 % 2 == 0) z1[R1 / 2] = scale(conj(x1[R1 / 2]), T(2));   // Z'[Nc/2] = 2 conj X[Nc/2]
#pragma unroll
    for (int j = 1; 2 * j < R1; ++j) real_pre_regs<T>(x1[j], x1[R1 - j], ldtab(twr + M * j), &z1[j], &z1[R1 - j]);
    radix_stage1_store<T, R1, R2, R3, +1>(0, z1, tw, buf);
    return;
  }
  if (2 * p == M) {                                              // k = M/2 + M j  <->  M/2 + M (R1-1-j)
#pragma unroll
    for (int j = 0; 2 * j < R1 - 1; ++j) real_pre_regs<T>(x1[j], x1[R1 - 1 - j], ldtab(twr + p + M * j), &z1[j], &z1[R1 - 1 - j]);
    if (R1 % 2 == 1) z1[(R1 - 1) / 2] = scale(conj(x1[(R1 - 1) / 2]), T(2));     // self-paired: k = Nc/2
    radix_stage1_store<T, R1, R2, R3, +1>(p, z1, tw, buf);
    return;
  }
  cpx<T> x2[R1], z2[R1];
#pragma unroll
  for (int j = 0; j < R1; ++j) x2[j] = spec_get<Z, true>(ibase, b2 + j * M, N);
#pragma unroll
  for (int j = 0; j < R1; ++j) real_pre_regs<T>(x1[j], x2[R1 - 1 - j], ldtab(twr + p + M * j), &z1[j], &z2[R1 - 1 - j]);
  radix_stage1_store<T, R1, R2, R3, +1>(p, z1, tw, buf);
  radix_stage1_store<T, R1, R2, R3, +1>(b2, z2, tw, buf);
  (void)NC;
}
// Measured against a -DRADIX_NO_PAIRS=1 build (profiles/r02b_radix.md): +7 ... +17 % for a first radix <= 10 (real 5184: 0.35 ->
// 0.41) and for the float cores from 3840 points (real 7680 / 9600 / 10240: 0.33 / 0.30 / 0.34 -> 0.38 / 0.32 / 0.37); with a
// radix-16 first stage on the smaller cores (two butterflies = 64 registers of data, half the threads loading) and in double
// precision it LOSES 10-30 % -- those keep the pre-rotation pass through shared memory.
template <typename T, int R1, int R2, int R3, int LM, int SIGN> PF_HD constexpr bool radix_pairs_in_wanted() {
  return (LM == L_R_ORD || LM == L_R_Z) && SIGN > 0 && R3 > 1 && !RADIX_NO_PAIRS && sizeof(T) == 4 && R1 * R2 * R3 >= 1024 &&
         (R1 <= 10 || (R1 <= 16 && R1 * R2 * R3 >= 3840));
}

// which (core, mode) takes the pair form (kernel and CPU stepping harness agree through this one function)
template <typename T, int R1, int R2, int R3, int SM, int SIGN> PF_HD constexpr bool radix_pairs_wanted() {
  return (SM == S_R_ORD || SM == S_R_Z) && SIGN < 0 && R3 > 1 && R3 <= 16 && !RADIX_NO_PAIRS &&
         R1 * R2 * R3 >= (sizeof(T) == 8 ? 256 : 1024);
}

#ifdef __CUDACC__
// TPC transforms per CTA iteration, TT threads each (blockDim.x = TPC * TT)
template <typename T, int R1, int R2, int R3, int LM, int SM, int SIGN, int TPC, int MINB>
__global__ void __launch_bounds__(TPC * RadixShape<R1, R2, R3>::TT, MINB)
k_cta_radix(const T* __restrict__ in, T* __restrict__ out, long long batch, const cpx<T>* tw, const cpx<T>* twr) {
  using S = RadixShape<R1, R2, R3>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  constexpr bool kReal = (LM == L_R_TIME || LM == L_R_ORD || LM == L_R_Z);
  constexpr int N = kReal ? 2 * S::NC : S::NC;
  constexpr bool partner = (SM == S_R_ORD || SM == S_R_Z);
  // forward real, three-stage cores with a last radix <= 16: the pair rotation runs in registers (radix_last_pairs).
  // Measured (profiles/r02b_radix.md): real 4608 ... 20480: +3 ... +30 % (5120: 0.42 -> 0.53, 9600: 0.36 -> 0.47); the small
  // two-stage cores LOSE 17-20 % (real 96: 0.54 -> 0.43: half of very few threads idle in the last stage) and keep the pairing pass.
  constexpr bool kPairs = radix_pairs_wanted<T, R1, R2, R3, SM, SIGN>();
  const int tid = threadIdx.x;
  const int tl = tid / S::TT;
  int li = tid - tl * S::TT;
  cpx<T>* buf = reinterpret_cast<cpx<T>*>(pf_smem_raw) + (size_t)tl * S::NCP;
  for (long long t0 = (long long)blockIdx.x * TPC; t0 < batch; t0 += (long long)gridDim.x * TPC) {
    asm volatile("" : "+r"(li), "+l"(tw), "+l"(twr));              // keep index math and table reads inside the loop (no spills)
    const bool live = t0 + tl < batch;
    const T* ibase = in + (t0 + tl) * (2LL * S::NC);
    T* obase = out + (t0 + tl) * (2LL * S::NC);
    constexpr bool prerot = (LM == L_R_ORD || LM == L_R_Z);       // backward real: pair-wise pre-rotation through `buf`
    cpx<T> v1[R1];
    constexpr bool kPairsIn = radix_pairs_in_wanted<T, R1, R2, R3, LM, SIGN>();
    if (kPairsIn) {                                               // backward real, large three-stage cores: pre-rotation in registers
      if (live) radix_first_pairs<T, R1, R2, R3, LM>(li, ibase, N, twr, tw, buf);
    } else if (prerot) {
      if (live) radix_prerotate<T, LM>(li, S::TT, ibase, N, S::NC, twr, buf);
      __syncthreads();
      if (live) radix_stage1_load<T, R1, R2, R3, LM, true>(li, ibase, N, twr, buf, v1);
      if (S::STAGES > 1) __syncthreads();                         // stage 1 overwrites the buffer it has just read
    } else if (live) {
      radix_stage1_load<T, R1, R2, R3, LM, false>(li, ibase, N, twr, buf, v1);
    }
    if (live && !kPairsIn) radix_stage1_store<T, R1, R2, R3, SIGN>(li, v1, tw, buf);
    if (S::STAGES == 1) {
      if (live) radix_emit<T, R1, S::M1, SM>(li, v1, obase, N, buf);
    } else if (S::STAGES == 2 && kPairs) {
      __syncthreads();
      if (live) radix_last_pairs<T, R2, S::M2, S::NC, SM, true, R1>(li, buf, obase, N, twr);
    } else if (S::STAGES == 2) {
      __syncthreads();
      cpx<T> a[R2];
      if (live) radix_stage2_read<T, R1, R2, R3>(li, buf, a);
      if (partner) __syncthreads();                               // `buf` becomes the natural-order buffer
      if (live && li < S::M2) { dft_small<R2, SIGN>(a); radix_emit<T, R2, S::M2, SM>(li, a, obase, N, buf); }
    } else {
      __syncthreads();
      {
        cpx<T> a[R2];
        if (live) radix_stage2_read<T, R1, R2, R3>(li, buf, a);
        __syncthreads();
        if (live) radix_stage2_write<T, R1, R2, R3, SIGN>(li, a, tw, buf);
      }
      __syncthreads();
      if (kPairs) {
        if (live) radix_last_pairs<T, R3, S::M3, S::NC, SM, false, R1>(li, buf, obase, N, twr);
      } else {
        cpx<T> c[R3];
        if (live) radix_stage3_read<T, R1, R2, R3>(li, buf, c);
        if (partner) __syncthreads();
        if (live && li < S::M3) { dft_small<R3, SIGN>(c); radix_emit<T, R3, S::M3, SM>(li, c, obase, N, buf); }
      }
    }
    if (partner && !kPairs) {
      __syncthreads();
      // bins k and Nc-k share their sum, difference and twiddle: one pass over half the spectrum (k = 0 also emits Nc/2)
      if (live) for (int k = li; k < S::NC / 2; k += S::TT) real_post_pair<SM, T>(obase, buf, k, N, S::NC, twr);
    }
    __syncthreads();                                              // buffer is rewritten by the next group's stage 1
  }
}
#endif  // __CUDACC__

}  // namespace pf
