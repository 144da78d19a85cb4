// cluster_kernels.cuh -- large complex cores (16384 .. 65536 points) with ONE HBM read and ONE HBM write per transform:
// a thread-block CLUSTER owns a transform, the rows live in the distributed shared memory of its CTAs.
//
//   Nc = R x N2,  R = CL x Q,  N2 = 256*C (the 16x16xC CTA core of cta_kernels.cuh),  cluster of CL CTAs, 16*C threads each
//   decimation in time:  X[k2 + N2*k1] = sum_n1 W_R^{n1 k1} * ( W_Nc^{n1 k2} * Y_n1[k2] ),   Y_n1 = FFT_N2( x[n1 + R*n2] )
//
// CTA `rank` transforms the rows n1 = rank + CL*q (q < Q) with the three register passes of k2_pass1/2/3.  Instead of
// writing a finished row to HBM (the two-pass plan of cta_hooks.cuh) every thread multiplies its 16 row outputs by
// W_Nc^{n1 k2} and stores them into the shared memory of the CTA that OWNS that k2 range
//        owner(k2) = k2 / S,   S = N2 / CL,   park[n1][k2 - owner*S]      (st through a mapa-translated address = DSMEM)
// -- 256-byte coalesced remote stores per warp.  After a cluster barrier each CTA finishes its S columns with radix-R
// register DFTs read from LOCAL shared memory and stores X in natural order (coalesced).  Barriers per transform: one
// split "park buffers free" (arrive after the combine, wait just before the first remote store of the next transform)
// and one "park buffers full".
//
// SCATTER (Q == 1): the rows are also DISTRIBUTED through DSMEM instead of being read with element stride R from L2
// (which moves 32-byte sectors for 8 useful bytes): CTA `rank` reads the contiguous slice x[rank*N2 .. (rank+1)*N2) with
// 128-bit loads, R consecutive points per thread, and sends point j of each run to CTA j (again 256 contiguous bytes per
// warp and destination); pass 1 then reads its row from local shared memory.  HBM and L2 only ever see dense traffic.
//
// Measured and dropped: the decimation-in-frequency order on a cluster (dense column reads -> radix-R in registers -> ONE DSMEM
// exchange -> rows from local shared memory -> element-stride-R stores): 0.29 / 0.27 / 0.18 of HBM peak at 16384 / 32768 /
// 65536 -- partial-sector stores are worse than the strided loads of the order above (0.41 / 0.30 / 0.23).
//
// Replaces, for these sizes, cfftf1_ps with its passf2/passf4 sweeps + finalize + zreorder (ref
// src/pffft_priv_impl.h:1004-1048, :122-251, :1195-1237, :1158-1193): N/4-point passes over a 128..512 KiB vector
// become one on-chip transform.
#pragma once
#include "cta_kernels.cuh"
#include "fastconv_kernels.cuh"   // k2_pass1_smem

namespace pf {

template <int C, int CL, int Q> struct KCL {
  using K = K2<C>;
  static constexpr int N2 = K::NC;
  static constexpr int R = CL * Q;
  static constexpr int S = N2 / CL;              // columns (k2 values) finished by one CTA
  static constexpr int NC = R * N2;
  static_assert(N2 % CL == 0 && S % K::T == 0, "every thread finishes a whole number of columns");
  static constexpr int COLS_PER_THREAD = S / K::T;
};

// ---- pass 1 in two halves, so that the loads of the NEXT row are in flight while the current one is parked and combined.
// cl_row_issue: thread m copies its 16 points x[n1 + R*(m + 16C*n_a)] with 8-byte cp.async (LDGSTS, no registers held)
// into the 16 tile slots it will itself overwrite in pass 1 -- thread-private slots, so no barrier is needed between the
// copy and its consumption, only cp.async.wait_all.  Register p of the radix-16 network holds point n_a = brev4(p).
template <int C, int R, typename T>
PF_HD void cl_row_issue(int m, const cpx<T>* row /* x + n1 */, cpx<T>* tile) {
  using K = K2<C>;
  const int jb = m / C, jc = m % C;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const cpx<T>* g = row + (long long)(m + K::BC * brev4(p)) * R;
    cpx<T>* d = tile + K::idx(p, jb, jc);
#ifdef __CUDA_ARCH__
    if constexpr (sizeof(cpx<T>) == 8)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(g) : "memory");
    else
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(d)), "l"(g) : "memory");
#else
    *d = *g;
#endif
  }
}
PF_HD void cl_row_landed() {
#ifdef __CUDA_ARCH__
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
// cl_row_pass1: radix-16 over n_a of the staged points, * W_N2^{m k_a}, back into the same slots (== k2_pass1 after its loads)
template <int C, int SIGN, typename T>
PF_HD void cl_row_pass1(int m, const cpx<T>* tw1, cpx<T>* tile) {
  using K = K2<C>;
  const int jb = m / C, jc = m % C;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = tile[K::idx(p, jb, jc)];
  reg_fft<16, SIGN>(v);
  tile[K::idx(0, jb, jc)] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[K::idx(ka, jb, jc)] = cmul_dir<SIGN>(v[ka], ldtab(tw1 + ka * K::BC + m));
}

// ---- park: thread t of the CTA that just transformed row n1 holds u[r*C + kc] = Y_n1[k2_out_index(t, r, kc)]
// remote(owner) -> base of the park buffer [R][S] of CTA `owner`
// twP: row-major twiddle table twP[n1*N2 + k2] = exp(-2 pi i n1 k2 / Nc): consecutive threads read consecutive entries
// (the natural table exp(-2 pi i k / Nc) would be read with stride n1: one 32-byte sector per 8 useful bytes)
template <int C, int CL, int Q, int SIGN, typename T, typename Remote>
PF_HD void cl_park(int t, int n1, const cpx<T> (&u)[16], const cpx<T>* twP, Remote remote) {
  using G = KCL<C, CL, Q>;
#pragma unroll
  for (int r = 0; r < 16 / C; ++r)
#pragma unroll
    for (int kc = 0; kc < C; ++kc) {
      const int k2 = k2_out_index<C>(t, r, kc);
      const int owner = k2 / G::S, j = k2 - owner * G::S;
      const cpx<T> v = (n1 == 0) ? u[r * C + kc] : cmul_dir<SIGN>(u[r * C + kc], ldtab(twP + n1 * G::N2 + k2));
      remote(owner)[n1 * G::S + j] = v;
    }
}

// ---- combine: CTA `rank` finishes columns k2 = rank*S + j from its own park buffer
template <int C, int CL, int Q, int SIGN, typename T>
PF_HD void cl_combine(int t, int rank, const cpx<T>* park, cpx<T>* dst) {
  using G = KCL<C, CL, Q>;
#pragma unroll 1
  for (int i = 0; i < G::COLS_PER_THREAD; ++i) {
    const int j = t + G::K::T * i;
    cpx<T> v[G::R];
#pragma unroll
    for (int n1 = 0; n1 < G::R; ++n1) v[n1] = park[n1 * G::S + j];
    dft_small<G::R, SIGN>(v);
    const int k2 = rank * G::S + j;
#pragma unroll
    for (int k1 = 0; k1 < G::R; ++k1) dst[k2 + G::N2 * k1] = v[k1];
  }
}

// ---- scatter (Q == 1, R == CL): run g of CTA `rank` = points x[rank*N2 + R*g + j], j < R  ->  row j, position rank*S + g
template <int R, typename T> PF_HD void load_run(const cpx<T>* p, cpx<T> (&e)[R]) {
#ifdef __CUDA_ARCH__
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
      const float4 v = reinterpret_cast<const float4*>(p)[j];
      e[2 * j] = mk<T>(v.x, v.y); e[2 * j + 1] = mk<T>(v.z, v.w);
    }
    return;
  }
#endif
#pragma unroll
  for (int j = 0; j < R; ++j) e[j] = p[j];
}
template <int C, int CL, typename T, typename Remote>
PF_HD void cl_scatter(int t, int rank, const cpx<T>* src, Remote remote) {
  using G = KCL<C, CL, 1>;
#pragma unroll
  for (int i = 0; i < G::COLS_PER_THREAD; ++i) {
    const int g = t + G::K::T * i;                 // g < S runs per CTA
    cpx<T> e[CL];
    load_run<CL, T>(src + (long long)rank * G::N2 + CL * g, e);
#pragma unroll
    for (int j = 0; j < CL; ++j) remote(j)[rank * G::S + g] = e[j];
  }
}

#ifdef __CUDACC__
// release: the DSMEM stores issued before it are visible to the peers after their wait.  The fence behind it also
// waits for every other outstanding store of the thread (the combine's global stores: ncu showed it as the second
// largest stall), so barriers that only announce "I have finished READING" use the relaxed form -- the loads they
// cover have returned, their values were consumed by the stores issued before the arrive.
PF_D void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
PF_D void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
PF_D void cluster_wait()   { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PF_D unsigned cluster_cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// generic address of `p` (a shared-memory variable of this CTA) in the shared memory of CTA `rank` of the cluster
template <typename P> PF_D P* cluster_map(P* p, unsigned rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
  return reinterpret_cast<P*>(out);
}

// park buffer of a peer CTA (generic address into distributed shared memory)
template <typename T> struct ClusterRemote {
  cpx<T>* park;
  PF_HD cpx<T>* operator()(int owner) const {
#ifdef __CUDA_ARCH__
    return cluster_map(park, (unsigned)owner);
#else
    return nullptr;                                             // device-only type; the CPU stepping harness has its own
#endif
  }
};

// gridDim.x = (#clusters) * CL, cluster dimension CL (launch attribute), blockDim.x = 16*C.
// shared memory: tile [N2] + park [Q*N2] complex words (SCATTER: the park buffer doubles as the row staging buffer).
// MODE 0: next row staged by cp.async (LDGSTS) into the thread's own tile slots;  1: rows distributed through DSMEM (SCATTER).
// (Measured and dropped: plain loads inside pass 1 with the next row only prefetched into L2 by prefetch.global.L2 --
//  0.38 / 0.27 / 0.16 of HBM peak at 16384 / 32768 / 65536 against 0.41 / 0.33 / 0.23 for the cp.async staging.)
template <typename T, int C, int CL, int Q, int SIGN, int MODE, int MINB>
__global__ void __launch_bounds__(16 * C, MINB)
k_cluster_fft(const T* in, T* out, long long batch, const cpx<T>* tw1, const cpx<T>* tw2, const cpx<T>* twP) {
  using K = K2<C>;
  using G = KCL<C, CL, Q>;
  constexpr bool SCATTER = MODE == 1;
  static_assert(!SCATTER || Q == 1, "the staging buffer aliases the park buffer: one row per CTA");
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* park = tile + G::N2;                                  // [R][S]  (== [Q][N2] words)
  const int t = threadIdx.x;
  const int rank = (int)cluster_cta_rank();
  const long long nclusters = gridDim.x / CL, cid = blockIdx.x / CL;
  const ClusterRemote<T> remote{park};
  cluster_arrive_relaxed();                                     // "park buffers free", phase 0
  if (!SCATTER && cid < batch)
    cl_row_issue<C, G::R, T>(t, reinterpret_cast<const cpx<T>*>(in) + cid * (long long)G::NC + rank, tile);
  for (long long tr = cid; tr < batch; tr += nclusters) {
    asm volatile("" : "+l"(tw1), "+l"(tw2), "+l"(twP));         // keep table reads in the loop (see cta_kernels.cuh)
    const cpx<T>* src = reinterpret_cast<const cpx<T>*>(in) + tr * (long long)G::NC;
    cpx<T>* dst = reinterpret_cast<cpx<T>*>(out) + tr * (long long)G::NC;
    if (SCATTER) {
      cluster_wait();                                           // every CTA finished its previous combine: staging free
      cl_scatter<C, CL, T>(t, rank, src, remote);
      cluster_arrive(); cluster_wait();                         // rows complete in every CTA
      k2_pass1_smem<C, SIGN, T>(t, park, tw1, tile);
      __syncthreads();
      cluster_arrive_relaxed();                                 // this CTA no longer reads its staging buffer
      k2_pass2<C, SIGN, T>(t, tw2, tile);
      __syncthreads();
      cpx<T> u[16];
      k2_pass3<C, SIGN, T>(t, tile, u);
      cluster_wait();                                           // nobody reads staging any more: it becomes the park buffer
      cl_park<C, CL, Q, SIGN, T>(t, rank, u, twP, remote);
    } else {
#pragma unroll 1
      for (int q = 0; q < Q; ++q) {
        const int n1 = rank + CL * q;
        cl_row_landed();                                        // own 16 points are in own tile slots
        cl_row_pass1<C, SIGN, T>(t, tw1, tile);
        __syncthreads();
        k2_pass2<C, SIGN, T>(t, tw2, tile);
        __syncthreads();
        cpx<T> u[16];
        k2_pass3<C, SIGN, T>(t, tile, u);
        __syncthreads();                                        // every pass-3 read done: the tile can take the next row
        if (q + 1 < Q) cl_row_issue<C, G::R, T>(t, src + n1 + CL, tile);
        else if (tr + nclusters < batch) cl_row_issue<C, G::R, T>(t, src + nclusters * (long long)G::NC + rank, tile);
        if (q == 0) cluster_wait();                             // every CTA finished its previous combine: park free
        cl_park<C, CL, Q, SIGN, T>(t, n1, u, twP, remote);
      }
    }
    cluster_arrive(); cluster_wait();                           // all rows parked (release/acquire orders the DSMEM stores)
    cl_combine<C, CL, Q, SIGN, T>(t, rank, park, dst);
    cluster_arrive_relaxed();                                   // "park buffers free" for the next transform
  }
  cluster_wait();                                               // no CTA exits while a peer may still address its memory
}

#endif  // __CUDACC__

}  // namespace pf
