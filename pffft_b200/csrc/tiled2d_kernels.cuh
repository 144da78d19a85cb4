// tiled2d_kernels.cuh -- large complex cores as a TILED two-dimensional transform: Nc = N1 x N2, N1 = 16*A1, N2 = 16*A2
// (A1, A2 in {8, 16}: 16384 = 128 x 128, 32768 = 256 x 128, 65536 = 256 x 256), both passes moving 128-byte runs.
//
//   n = n1 + N1*n2,  k = k2 + N2*k1:
//   X[k2 + N2*k1] = sum_n1 W_N1^{n1 k1} * W_Nc^{n1 k2} * ( sum_n2 x[n1 + N1*n2] W_N2^{n2 k2} )
//
//   pass A (k_t2d_A): one CTA = 16 neighbouring columns n1 = 16c .. 16c+15 of one transform (16 x N2 points, 16 per thread).
//        Loads: for every n2 the 16 columns are ONE 128-byte run.  N2-point FFTs down the columns (radix 16 in registers,
//        exchange through shared memory, radix A2), * W_Nc^{n1 k2}, stored TRANSPOSED: S[k2][n1] -- again 128-byte runs.
//   pass C (k_t2d_C): one CTA = 16 neighbouring rows k2 = 16d .. 16d+15 of S (16 x N1 points): N1-point FFTs along the rows
//        (radix 16, exchange with an XOR swizzle, radix A1), stored as X[k2 + N2*k1]: the 16 rows are one 128-byte run per k1.
//
// Every HBM access of both passes is a full 128-byte line (the split plans of cta_hooks.cuh read rows with element stride
// R: one 32-byte sector per 8 useful bytes, LSU bound at R = 16 -- profiles/r01b_large_n.md).  Two HBM round trips
// (ceiling 0.5 of the roofline); the same two phase pairs are what a cluster version exchanges through DSMEM instead of S.
//
// Measured on B200 (first hardware run, untuned; profiles/r01b_large_n.md): 16384: 0.39, 32768: 0.42, 65536: 0.41 forward /
// 0.44 backward of the HBM peak (split plan: 0.38 / 0.34 / 0.26).  Index algebra also verified by CPU stepping
// (tests/test_host_logic.py via tests/emu).  80 registers (a 64-register build spills 16-24 of them): tuning left for round 2.
//
// Replaces, for these sizes, the cfftf1_ps sweeps + finalize + zreorder of the reference (src/pffft_priv_impl.h:1004-1048,
// :122-251, :1195-1237, :1158-1193).
#pragma once
#include "butterfly.cuh"
#include "cta_kernels.cuh"   // brev4, ldtab
#include "plan.h"          // unit_root (host table fill)
#include "cluster_kernels.cuh"   // cluster_arrive / cluster_wait / ClusterRemote (cluster-fused form)

namespace pf {

template <int A> PF_HD constexpr int brevA(int p) { return ct::bitrev(p, ct::ilog2(A)); }

// radix-A register FFTs on the PP = 16/A groups u[r*A .. r*A + A) (bit-reversed in, natural out)
template <int A, int SIGN, typename T> PF_HD void t2d_small_ffts(cpx<T> (&u)[16]) {
  if constexpr (A == 16) dit_fft<16, SIGN, 0, 1>(u);
  else { static_assert(A == 8, "A in {8, 16}"); dit_fft<8, SIGN, 0, 1>(u); dit_fft<8, SIGN, 8, 1>(u); }
}

// table layouts (all exp(-2 pi i .)):
//   twA [k_a*A2 + q]                     = W_N2^{q k_a}                 (N2 entries)
//   twC [k1a*A1 + q1]                    = W_N1^{q1 k1a}                (N1 entries)
//   tw2d[((c*A2 + k_b)*16 + k_a)*16 + j] = W_Nc^{(16c + j)(k_a + 16 k_b)}  (Nc entries; lanes j read consecutive entries)
template <int A1, int A2> struct T2D {
  static constexpr int N1 = 16 * A1, N2 = 16 * A2, NC = N1 * N2;
  static constexpr int TA = 16 * A2, TC = 16 * A1;           // threads of pass A / pass C
  PF_HD static int idx2d(int c, int k_a, int k_b, int j) { return ((c * A2 + k_b) * 16 + k_a) * 16 + j; }
};

// ---- pass A, first half: thread t = j + 16*q loads column n1 = 16c + j at n2 = q + A2*i (i < 16), radix 16 over i -> k_a,
//      * W_N2^{q k_a}, into tile[(k_a*A2 + q)*16 + j]   (lanes j -> consecutive words: conflict free both ways)
template <int A1, int A2, int SIGN, typename T>
PF_HD void t2d_A1(int t, const cpx<T>* cols /* x + 16c */, const cpx<T>* twA, cpx<T>* tile) {
  using G = T2D<A1, A2>;
  const int j = t & 15, q = t >> 4;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = cols[j + G::N1 * (q + A2 * brev4(p))];
  reg_fft<16, SIGN>(v);
  tile[(0 * A2 + q) * 16 + j] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[(ka * A2 + q) * 16 + j] = cmul_dir<SIGN>(v[ka], ldtab(twA + ka * A2 + q));
}
// ---- pass A, second half: thread t = j + 16*g finishes k_a = g + A2*r (r < 16/A2): radix A2 over q -> k_b,
//      k2 = k_a + 16*k_b, * W_Nc^{n1 k2}, to the transposed scratch S[k2*N1 + 16c + j]
// Sink: row(k2) -> where row k2 of the transposed intermediate lives (element n1 at row(k2)[n1])
template <typename T, int N1> struct T2DScratchSink {           // the two-pass plan: S[k2][n1] in global memory
  cpx<T>* S;
  PF_HD cpx<T>* row(int k2) const { return S + (long long)k2 * N1; }
};
template <int A1, int A2, int SIGN, typename T, typename Sink>
PF_HD void t2d_A2(int t, int c, const cpx<T>* tile, const cpx<T>* tw2d, Sink sink) {
  using G = T2D<A1, A2>;
  const int j = t & 15, g = t >> 4;
  cpx<T> u[16];
#pragma unroll
  for (int r = 0; r < 16 / A2; ++r)
#pragma unroll
    for (int p = 0; p < A2; ++p) u[r * A2 + p] = tile[((g + A2 * r) * A2 + brevA<A2>(p)) * 16 + j];
  t2d_small_ffts<A2, SIGN>(u);
#pragma unroll
  for (int r = 0; r < 16 / A2; ++r)
#pragma unroll
    for (int kb = 0; kb < A2; ++kb) {
      const int ka = g + A2 * r, k2 = ka + 16 * kb;
      const cpx<T> w = ldtab(tw2d + G::idx2d(c, ka, kb, j));
      sink.row(k2)[16 * c + j] = cmul_dir<SIGN>(u[r * A2 + kb], w);
    }
}
// ---- pass C, first half: thread t = q1 + A1*k2' loads row k2 = 16d + k2' at n1 = q1 + A1*i (i < 16), radix 16 over i -> k1a,
//      * W_N1^{q1 k1a}, into tile[(k1a*A1 + q1)*16 + (k2' ^ swz(q1))]
template <int A1> PF_HD int t2d_swz(int q1) { return (q1 * (16 / A1)) & 15; }
template <int A1, int A2, int SIGN, typename T>
PF_HD void t2d_C1(int t, const cpx<T>* rows /* S + 16d*N1 */, const cpx<T>* twC, cpx<T>* tile) {
  using G = T2D<A1, A2>;
  const int q1 = t % A1, k2p = t / A1;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = rows[k2p * G::N1 + q1 + A1 * brev4(p)];
  reg_fft<16, SIGN>(v);
  const int col = k2p ^ t2d_swz<A1>(q1);
  tile[(0 * A1 + q1) * 16 + col] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[(ka * A1 + q1) * 16 + col] = cmul_dir<SIGN>(v[ka], ldtab(twC + ka * A1 + q1));
}
// ---- pass C, second half: thread t = k2' + 16*g finishes k1a = g + A1*r: radix A1 over q1 -> k1b, k1 = k1a + 16*k1b,
//      X[(16d + k2') + N2*k1]   (lanes k2' -> one 128-byte run per k1)
template <int A1, int A2, int SIGN, typename T>
PF_HD void t2d_C2(int t, const cpx<T>* tile, cpx<T>* xblk /* X + 16d */) {
  using G = T2D<A1, A2>;
  const int k2p = t & 15, g = t >> 4;
  cpx<T> u[16];
#pragma unroll
  for (int r = 0; r < 16 / A1; ++r)
#pragma unroll
    for (int p = 0; p < A1; ++p) {
      const int qq = brevA<A1>(p);
      u[r * A1 + p] = tile[((g + A1 * r) * A1 + qq) * 16 + (k2p ^ t2d_swz<A1>(qq))];
    }
  t2d_small_ffts<A1, SIGN>(u);
#pragma unroll
  for (int r = 0; r < 16 / A1; ++r)
#pragma unroll
    for (int kb = 0; kb < A1; ++kb) xblk[k2p + (long long)G::N2 * ((g + A1 * r) + 16 * kb)] = u[r * A1 + kb];
}

#ifdef __CUDACC__
// grid-stride over the tiles (transform b, column block c): tile index = b * (N1/16) + c
template <typename T, int A1, int A2, int SIGN, int MINB>
__global__ void __launch_bounds__(16 * A2, MINB)
k_t2d_A(const cpx<T>* __restrict__ x, cpx<T>* __restrict__ S, long long batch, const cpx<T>* twA, const cpx<T>* tw2d) {
  using G = T2D<A1, A2>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);      // [16][A2][16] = 16*N2 words
  const int t = threadIdx.x;
  constexpr int CB = G::N1 / 16;
  const long long tiles = batch * CB;
  for (long long w = blockIdx.x; w < tiles; w += gridDim.x) {
    asm volatile("" : "+l"(twA), "+l"(tw2d));
    const long long b = w / CB;
    const int c = (int)(w - b * CB);
    t2d_A1<A1, A2, SIGN, T>(t, x + b * (long long)G::NC + 16 * c, twA, tile);
    __syncthreads();
    t2d_A2<A1, A2, SIGN, T>(t, c, tile, tw2d, T2DScratchSink<T, G::N1>{S + b * (long long)G::NC});
    __syncthreads();
  }
}
template <typename T, int A1, int A2, int SIGN, int MINB>
__global__ void __launch_bounds__(16 * A1, MINB)
k_t2d_C(const cpx<T>* __restrict__ S, cpx<T>* __restrict__ X, long long batch, const cpx<T>* twC) {
  using G = T2D<A1, A2>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);      // [16][A1][16] = 16*N1 words
  const int t = threadIdx.x;
  constexpr int RB = G::N2 / 16;
  const long long tiles = batch * RB;
  for (long long w = blockIdx.x; w < tiles; w += gridDim.x) {
    asm volatile("" : "+l"(twC));
    const long long b = w / RB;
    const int d = (int)(w - b * RB);
    t2d_C1<A1, A2, SIGN, T>(t, S + b * (long long)G::NC + (long long)16 * d * G::N1, twC, tile);
    __syncthreads();
    t2d_C2<A1, A2, SIGN, T>(t, tile, X + b * (long long)G::NC + 16 * d);
    __syncthreads();
  }
}
#endif  // __CUDACC__

// =====================================================================================================================
// CLUSTER-FUSED form: the same two phase pairs, but pass A hands its rows to the CTA that runs pass C through DISTRIBUTED
// SHARED MEMORY instead of the scratch array -- one HBM read and one HBM write per transform (ceiling 1.0 instead of 0.5).
//   cluster of CL CTAs per transform; column blocks c = rank + CL*qa (qa < QA = N1/16/CL) are transformed one after the
//   other; value (n1, k2) goes to row block db = k2/16, owned by CTA db % CL as its block number db / CL:
//        park[(db / CL)][k2 % 16][n1]          (lanes j -> 128-byte runs, as in the scratch version)
//   one cluster barrier; then every CTA runs pass C for its QC = N2/16/CL row blocks from LOCAL shared memory.
//   Barrier protocol as in cluster_kernels.cuh: "parks free" (relaxed arrive after the last pass-C1 read, wait before the
//   first remote store of the next transform) and "parks full" (release / acquire).
// STATUS (end of round 1): arithmetic and index algebra verified by CPU stepping (tests/test_host_logic.py); NOT YET RUN ON
// HARDWARE.  Opt-in only (PFFFT_B200_TILED2D=2); its GPU tests are gated (PFFFT_B200_TEST_T2D_CLUSTER=1).
// =====================================================================================================================
template <int A1, int A2, int CL> struct T2DC {
  using G = T2D<A1, A2>;
  static constexpr int CB = G::N1 / 16, RB = G::N2 / 16;       // column blocks (pass A) / row blocks (pass C) per transform
  static_assert(CB % CL == 0 && RB % CL == 0, "blocks divide evenly over the cluster");
  static constexpr int QA = CB / CL, QC = RB / CL;
  static constexpr int NT = G::TA > G::TC ? G::TA : G::TC;       // threads per CTA
  static constexpr int TILE = 16 * (G::N1 > G::N2 ? G::N1 : G::N2);
  static constexpr int PARK_BLOCK = 16 * G::N1;                  // words of one row block [16][N1]
  static constexpr size_t kSmem = (size_t)(TILE + QC * PARK_BLOCK);   // in complex words
};
// Remote: rank -> base of that CTA's park buffer
template <typename T, int A1, int A2, int CL, typename Remote> struct T2DClusterSink {
  Remote remote;
  PF_HD cpx<T>* row(int k2) const {
    const int db = k2 >> 4;
    return remote(db % CL) + (db / CL) * T2DC<A1, A2, CL>::PARK_BLOCK + (k2 & 15) * T2D<A1, A2>::N1;
  }
};

#ifdef __CUDACC__
template <typename T, int A1, int A2, int CL, int SIGN, int MINB>
__global__ void __launch_bounds__((T2DC<A1, A2, CL>::NT), MINB)
k_t2d_cluster(const cpx<T>* __restrict__ x, cpx<T>* __restrict__ X, long long batch,
              const cpx<T>* twA, const cpx<T>* twC, const cpx<T>* tw2d) {
  using G = T2D<A1, A2>;
  using K = T2DC<A1, A2, CL>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* park = tile + K::TILE;                                // [QC][16][N1]
  const int t = threadIdx.x;
  const int rank = (int)cluster_cta_rank();
  const long long nclusters = gridDim.x / CL, cid = blockIdx.x / CL;
  const T2DClusterSink<T, A1, A2, CL, ClusterRemote<T>> sink{ClusterRemote<T>{park}};
  cluster_arrive_relaxed();                                     // "parks free", phase 0
  for (long long tr = cid; tr < batch; tr += nclusters) {
    asm volatile("" : "+l"(twA), "+l"(twC), "+l"(tw2d));
    const cpx<T>* src = x + tr * (long long)G::NC;
    cpx<T>* dst = X + tr * (long long)G::NC;
#pragma unroll 1
    for (int qa = 0; qa < K::QA; ++qa) {
      const int c = rank + CL * qa;
      if (t < G::TA) t2d_A1<A1, A2, SIGN, T>(t, src + 16 * c, twA, tile);
      __syncthreads();
      if (qa == 0) cluster_wait();                              // every CTA has finished reading its parks
      if (t < G::TA) t2d_A2<A1, A2, SIGN, T>(t, c, tile, tw2d, sink);
      __syncthreads();                                          // tile free for the next column block / pass C
    }
    cluster_arrive(); cluster_wait();                           // parks full (release / acquire orders the DSMEM stores)
#pragma unroll 1
    for (int qc = 0; qc < K::QC; ++qc) {
      const int d = rank + CL * qc;
      if (t < G::TC) t2d_C1<A1, A2, SIGN, T>(t, park + qc * K::PARK_BLOCK, twC, tile);
      __syncthreads();
      if (qc == K::QC - 1) cluster_arrive_relaxed();            // last read of this CTA's parks is done
      if (t < G::TC) t2d_C2<A1, A2, SIGN, T>(t, tile, dst + 16 * d);
      __syncthreads();
    }
  }
  cluster_wait();                                               // no CTA exits while a peer may still address its memory
}
#endif  // __CUDACC__

// =====================================================================================================================
// GENERAL radices: N1 = 16*A1, N2 = 16*A2 with A1, A2 any size of the register DFT library (2,3,4,5,6,8,9,10,12,15,16), i.e.
// Nc = 256*A1*A2: 7680 = 96 x 80, 9216 = 96 x 96, 12288 = 128 x 96, 20480 = 160 x 128, 24576 = 192 x 128, 36864 = 192 x 192,
// 40960 = 256 x 160, 49152 = 256 x 192, 61440 = 256 x 240 ...  Same passes and tables; 256 threads per CTA:
//   first halves: 16*A threads hold 16 points each (pass C pads its rows to 16 lanes so the XOR swizzle stays conflict free);
//   second halves: every thread finishes ONE (column, k_a) pair with a radix-A register DFT (natural order in and out).
// STATUS (end of round 1): verified by CPU stepping for every shape below; NOT YET RUN ON HARDWARE.  Opt-in only
// (PFFFT_B200_TILED2D_GENERAL=1), GPU tests gated (PFFFT_B200_TEST_T2D_GENERAL=1).
// =====================================================================================================================
template <int A1, int A2, int SIGN, typename T, typename Sink>
PF_HD void t2dg_A2(int t, int c, const cpx<T>* tile, const cpx<T>* tw2d, Sink sink) {   // t < 256: j = t & 15, k_a = t >> 4
  using G = T2D<A1, A2>;
  const int j = t & 15, ka = t >> 4;
  cpx<T> u[A2];
#pragma unroll
  for (int q = 0; q < A2; ++q) u[q] = tile[(ka * A2 + q) * 16 + j];
  dft_small<A2, SIGN>(u);
#pragma unroll
  for (int kb = 0; kb < A2; ++kb)
    sink.row(ka + 16 * kb)[16 * c + j] = cmul_dir<SIGN>(u[kb], ldtab(tw2d + G::idx2d(c, ka, kb, j)));
}
// pass C, first half with rows padded to 16 lanes: thread t = q1 + 16*k2' works when q1 < A1
template <int A1, int A2, int SIGN, typename T>
PF_HD void t2dg_C1(int t, const cpx<T>* rows, const cpx<T>* twC, cpx<T>* tile) {
  using G = T2D<A1, A2>;
  const int q1 = t & 15, k2p = t >> 4;
  if (q1 >= A1) return;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = rows[k2p * G::N1 + q1 + A1 * brev4(p)];
  reg_fft<16, SIGN>(v);
  const int col = k2p ^ q1;
  tile[(0 * A1 + q1) * 16 + col] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[(ka * A1 + q1) * 16 + col] = cmul_dir<SIGN>(v[ka], ldtab(twC + ka * A1 + q1));
}
template <int A1, int A2, int SIGN, typename T>
PF_HD void t2dg_C2(int t, const cpx<T>* tile, cpx<T>* xblk) {        // t < 256: k2' = t & 15, k1a = t >> 4
  using G = T2D<A1, A2>;
  const int k2p = t & 15, ka = t >> 4;
  cpx<T> u[A1];
#pragma unroll
  for (int q = 0; q < A1; ++q) u[q] = tile[(ka * A1 + q) * 16 + (k2p ^ q)];
  dft_small<A1, SIGN>(u);
#pragma unroll
  for (int kb = 0; kb < A1; ++kb) xblk[k2p + (long long)G::N2 * (ka + 16 * kb)] = u[kb];
}

#ifdef __CUDACC__
template <typename T, int A1, int A2, int SIGN, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_t2dg_A(const cpx<T>* __restrict__ x, cpx<T>* __restrict__ S, long long batch, const cpx<T>* twA, const cpx<T>* tw2d) {
  using G = T2D<A1, A2>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);      // 16*N2 words
  const int t = threadIdx.x;
  constexpr int CB = G::N1 / 16;
  const long long tiles = batch * CB;
  for (long long w = blockIdx.x; w < tiles; w += gridDim.x) {
    asm volatile("" : "+l"(twA), "+l"(tw2d));
    const long long b = w / CB;
    const int c = (int)(w - b * CB);
    if (t < G::TA) t2d_A1<A1, A2, SIGN, T>(t, x + b * (long long)G::NC + 16 * c, twA, tile);
    __syncthreads();
    t2dg_A2<A1, A2, SIGN, T>(t, c, tile, tw2d, T2DScratchSink<T, G::N1>{S + b * (long long)G::NC});
    __syncthreads();
  }
}
template <typename T, int A1, int A2, int SIGN, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_t2dg_C(const cpx<T>* __restrict__ S, cpx<T>* __restrict__ X, long long batch, const cpx<T>* twC) {
  using G = T2D<A1, A2>;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);      // 16*N1 words
  const int t = threadIdx.x;
  constexpr int RB = G::N2 / 16;
  const long long tiles = batch * RB;
  for (long long w = blockIdx.x; w < tiles; w += gridDim.x) {
    asm volatile("" : "+l"(twC));
    const long long b = w / RB;
    const int d = (int)(w - b * RB);
    t2dg_C1<A1, A2, SIGN, T>(t, S + b * (long long)G::NC + (long long)16 * d * G::N1, twC, tile);
    __syncthreads();
    t2dg_C2<A1, A2, SIGN, T>(t, tile, X + b * (long long)G::NC + 16 * d);
    __syncthreads();
  }
}
#endif  // __CUDACC__

// host: fill [twA: N2][twC: N1][tw2d: Nc]
template <typename T, int A1, int A2> void t2d_fill_tables(T* dst) {
  using G = T2D<A1, A2>;
  long double c, s;
  T* a = dst;
  for (int ka = 0; ka < 16; ++ka) for (int q = 0; q < A2; ++q) { pfplan::unit_root((long long)q * ka, G::N2, &c, &s); a[2 * (ka * A2 + q)] = (T)c; a[2 * (ka * A2 + q) + 1] = (T)s; }
  T* cc = dst + 2 * (size_t)G::N2;
  for (int ka = 0; ka < 16; ++ka) for (int q1 = 0; q1 < A1; ++q1) { pfplan::unit_root((long long)q1 * ka, G::N1, &c, &s); cc[2 * (ka * A1 + q1)] = (T)c; cc[2 * (ka * A1 + q1) + 1] = (T)s; }
  T* d = cc + 2 * (size_t)G::N1;
  for (int cb = 0; cb < G::N1 / 16; ++cb) for (int kb = 0; kb < A2; ++kb) for (int ka = 0; ka < 16; ++ka) for (int j = 0; j < 16; ++j) {
    pfplan::unit_root((long long)(16 * cb + j) * (ka + 16 * kb), G::NC, &c, &s);
    const int i = G::idx2d(cb, ka, kb, j);
    d[2 * i] = (T)c; d[2 * i + 1] = (T)s;
  }
}

}  // namespace pf
