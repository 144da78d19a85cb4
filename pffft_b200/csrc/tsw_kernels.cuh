// tsw_kernels.cuh -- the tiled Stockham pipeline (ts_kernels.cuh) with WARP-SIZED work items: one warp owns a tile of a
// few neighbouring columns with all R = 16*A points of each, runs both halves of the radix-R pass on it and exchanges the
// data through a private shared-memory tile with __syncwarp() only.  No CTA barrier exists in the persistent loop.
//
// Why (profiles/r02b_large_n.md): the CTA-sized work items of ts_kernels.cuh spend their time at barriers -- after the
// packed-arithmetic build and the removal of the 64-bit modulo calls the 65536-point pipeline issues 0.42 of its slots and
// stalls 6.9 warps per issued instruction at `barrier` (ncu r02b_ncu_ts4_65536.txt): every one of the three barriers of a
// work item waits for the slowest of eight warps, and the loads of a CTA are all in flight at the same moment.  The
// N = 1024 kernel (fast_kernels.cuh), which this file copies in spirit, has no barrier and reaches the HBM roofline with
// 16 resident warps per SM.
//
// Shapes, for a pass of radix R = 16*A, A in {2, 4, 8, 16} (power-of-two cores 8192 ... 2^26; everything else stays on
// ts_kernels.cuh):   QL = min(A, 8) sub-sequences are spread over the lanes, U = A/QL (1 or 2) per lane,
//                    COLS = 32/QL columns per work item (4 for R = 128/256: 32-byte runs, sector exact).
//   phase 1: lane (c = lane % COLS, qh = lane / COLS), q = qh + QL*u: loads x[b0 + c + m*(q + A*i)], i < 16, radix-16 register
//            FFT over i -> k_a, * W_R^{q k_a} (shared-memory table), into the warp's tile at q*S + k_a*COLS + c, S = 17*COLS
//            (the odd multiple keeps the 64-bit accesses of both phases bank-conflict free: a half-warp covers COLS columns x
//            16/COLS values of q -- or of k_a in phase 2 -- and S mod 16 = COLS spreads them over the 16 bank pairs);
//   phase 2: lane (c, kh = lane / COLS), k_a = kh + QL*w, w < 16/QL: radix-A register DFT over q -> k_b, k = k_a + 16 k_b,
//            * W_Nc^{s p k}, stored exactly where ts_kernels.cuh stores it (first pass y[R*b + k], later y[q' + s*(R*p + k)]).
//            The LAST pass (s*R = Nc) takes W_R^{(p k) mod R} from a shared-memory table: no global twiddle loads.
// Work distribution, dependency counters, rings, lag: unchanged (ts_kernels.cuh), per WARP instead of per CTA: lane 0 reads
// the counters (relaxed, one item ahead), polls when it must, and signals completion with red.release after a
// __syncwarp() (bar.warp.sync orders the other lanes' stores before it, as bar.sync does for a CTA in ts_kernels.cuh).
#pragma once
#include "ts_kernels.cuh"

namespace pf {

enum { kTswWarps = 4 };                                       // warps per CTA

template <int A> struct TswShape {
  static constexpr int R = 16 * A;
  static constexpr int QL = A < 8 ? A : 8;                    // sub-sequences q spread over the lanes
  static constexpr int U = A / QL;                            // sub-sequences per lane
  static constexpr int COLS = 32 / QL;                        // columns per work item
  static constexpr int W = 16 / QL;                           // (column, k_a) pairs per lane in phase 2
  static constexpr int S = 17 * COLS;                         // tile stride between sub-sequences
  static constexpr int TILE = A * S;                          // entries of one warp's tile
  PF_HD static int idx(int q, int ka, int c) { return q * S + ka * COLS + c; }
};
PF_HD constexpr int tsw_cols_for(int A) { return 32 / (A < 8 ? A : 8); }
PF_HD constexpr bool tsw_radix_ok(int A) { return A == 2 || A == 4 || A == 8 || A == 16; }
constexpr int kTswTileMax = 16 * 17 * 4;                      // A = 16: 1088 entries (A = 2: 544, A = 4: 544, A = 8: 544)

// ---- phase 1
template <int A, int SIGN, typename T>
PF_HD void tsw_phase1(int lane, int b0, const cpx<T>* src /* transform base */, int m, const cpx<T>* twR /* [ka*A + q] */, cpx<T>* tile) {
  using S = TswShape<A>;
  const int c = lane % S::COLS, qh = lane / S::COLS;
  cpx<T> v[S::U][16];
#pragma unroll
  for (int u = 0; u < S::U; ++u) {
    const cpx<T>* p0 = src + b0 + c + (long long)m * (qh + S::QL * u);
#pragma unroll
    for (int p = 0; p < 16; ++p) v[u][p] = ld_l2(p0 + (long long)m * (A * brev4(p)));
  }
#pragma unroll
  for (int u = 0; u < S::U; ++u) {
    const int q = qh + S::QL * u;
    reg_fft<16, SIGN>(v[u]);
    tile[S::idx(q, 0, c)] = v[u][0];
#pragma unroll
    for (int ka = 1; ka < 16; ++ka) tile[S::idx(q, ka, c)] = cmul_dir<SIGN>(v[u][ka], twR[ka * A + q]);
  }
}
// ---- phase 2.  KIND 0: first pass (s = 1), 1: middle pass, 2: last pass (s*R = Nc; twL[e] = exp(-2 pi i e / R), e < R)
template <int A, int KIND, int SIGN, typename T>
PF_HD void tsw_phase2(int lane, int b0, int m, int s, const cpx<T>* tw, const cpx<T>* twL, const cpx<T>* tile, cpx<T>* dst) {
  using S = TswShape<A>;
  const int c = lane % S::COLS, kh = lane / S::COLS;
#pragma unroll
  for (int w = 0; w < S::W; ++w) {
    const int ka = kh + S::QL * w;
    cpx<T> u[A];
#pragma unroll
    for (int q = 0; q < A; ++q) u[q] = tile[S::idx(q, ka, c)];
    dft_small<A, SIGN>(u);
    if (KIND == 0) {                                          // y[R*b + k] = W_Nc^{b k} u,  W^{b k} = W^{b k_a} * W^{16 b k_b}
      const int b = b0 + c;
      const cpx<T> w1 = ldtab(tw + b * ka);
      cpx<T>* o = dst + (long long)S::R * b + ka;
      o[0] = cmul_dir<SIGN>(u[0], w1);
#pragma unroll
      for (int kb = 1; kb < A; ++kb) o[16 * kb] = cmul_dir<SIGN>(u[kb], cmul(w1, ldtab(tw + 16 * b * kb)));
    } else {
      const int p = b0 / s, e0 = p * s;                       // COLS divides s: p is uniform over the item
      cpx<T>* o = dst + (b0 - e0 + c) + (long long)s * ((long long)S::R * p + ka);
      const long long ks = 16LL * s;
      if (e0 == 0) {
#pragma unroll
        for (int kb = 0; kb < A; ++kb) o[ks * kb] = u[kb];
      } else if (KIND == 2) {                                 // W_Nc^{s p k} = W_R^{p k}: exponent-indexed shared table
        int e = (p * ka) & (S::R - 1);
        const int step = (16 * p) & (S::R - 1);
#pragma unroll
        for (int kb = 0; kb < A; ++kb) { o[ks * kb] = cmul_dir<SIGN>(u[kb], twL[e]); e = (e + step) & (S::R - 1); }
      } else {
#pragma unroll
        for (int kb = 0; kb < A; ++kb) o[ks * kb] = cmul_dir<SIGN>(u[kb], ldtab(tw + (long long)e0 * (ka + 16 * kb)));
      }
    }
  }
}

// one work item, phase by phase (__syncwarp between them in the kernel; tests/emu steps them lane by lane)
template <int A, int SIGN, typename T>
PF_HD void tsw_item_phase(int phase, int lane, int item, const TsStage& st, int Nc, const cpx<T>* src, cpx<T>* dst,
                          const cpx<T>* tw, const cpx<T>* twR, const cpx<T>* twL, cpx<T>* tile) {
  const int b0 = TswShape<A>::COLS * item;
  if (phase == 0) { tsw_phase1<A, SIGN, T>(lane, b0, src, st.m, twR + st.tw_off, tile); return; }
  if (st.kind == TS_FIRST) tsw_phase2<A, 0, SIGN, T>(lane, b0, st.m, st.s, tw, twL, tile, dst);
  else if ((long long)st.s * (16 * A) == Nc) tsw_phase2<A, 2, SIGN, T>(lane, b0, st.m, st.s, tw, twL, tile, dst);
  else tsw_phase2<A, 1, SIGN, T>(lane, b0, st.m, st.s, tw, twL, tile, dst);
}
template <int SIGN, typename T>
PF_HD void tsw_item_phase_any(int phase, int lane, int item, const TsStage& st, int Nc, const cpx<T>* src, cpx<T>* dst,
                              const cpx<T>* tw, const cpx<T>* twR, const cpx<T>* twL, cpx<T>* tile) {
  switch (st.A) {
    case 2:  tsw_item_phase<2, SIGN, T>(phase, lane, item, st, Nc, src, dst, tw, twR, twL, tile); break;
    case 4:  tsw_item_phase<4, SIGN, T>(phase, lane, item, st, Nc, src, dst, tw, twR, twL, tile); break;
    case 8:  tsw_item_phase<8, SIGN, T>(phase, lane, item, st, Nc, src, dst, tw, twR, twL, tile); break;
    case 16: tsw_item_phase<16, SIGN, T>(phase, lane, item, st, Nc, src, dst, tw, twR, twL, tile); break;
    default: break;
  }
}

#ifdef __CUDACC__
// THE PERSISTENT LOOP, one work item per warp and iteration.  Shared memory: [per-radix tables twR][last-pass table twL]
// [kTswWarps tiles].
template <typename T, int SIGN, int MINB>
__global__ void __launch_bounds__(kTswWarps * 32, MINB) k_tsw_pipeline(const __grid_constant__ TsParams<T> P, int twL_entries) {
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* twRs = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* twLs = twRs + ((P.twR_entries + 15) & ~15);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cpx<T>* tile = twLs + ((twL_entries + 15) & ~15) + warp * kTswTileMax;
  __shared__ TsStage ST[kTsMaxStages];
  for (int i = threadIdx.x; i < P.twR_entries; i += kTswWarps * 32) twRs[i] = P.twR[i];
  for (int i = threadIdx.x; i < twL_entries; i += kTswWarps * 32) twLs[i] = P.twR[P.twR_entries + i];   // (appended by the host)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kTsMaxStages; ++i) ST[i] = P.st[i];
  }
  __syncthreads();
  // stage-specialised workers (ts_worker): this warp serves one stage for the whole launch
  const TsWorker wk = ts_worker(P, ST, blockIdx.x * kTswWarps + warp, gridDim.x * kTswWarps);
  const int stage_i = wk.stage;
  const TsStage& st = ST[stage_i];
  const unsigned tiles = (unsigned)st.tiles;
  int cur_ready = 0;                                              // warp-uniform
  unsigned* pending = nullptr;                                    // completion signal of the previous item, not yet sent
  for (unsigned cur = wk.q0; cur < wk.count; cur += wk.step) {
    const unsigned nxt = cur + wk.step;
    const long long tr = (long long)(cur / tiles);
    const int item = (int)(cur - (unsigned)tr * tiles);
    // ---- readiness of item i (poll only if the early look did not already show it), early look at item i+1.  EVERY lane
    // runs this control code (same addresses: one request per warp instruction): the first version gave it to lane 0 alone,
    // and the warps ran the whole loop split in two halves of 16 lanes (ncu: 16 threads per executed instruction, every
    // __syncwarp on its divergent slow path, twice the instructions)
    const TsDeps d = ts_deps(P, ST, stage_i, tr);
    if (!cur_ready) {
      // nothing may be owed while this warp waits: others may be waiting for exactly that signal
      if (pending) { if (lane == 0) ts_red_release(pending); pending = nullptr; }
      for (;;) {
        const unsigned a = d.in_ctr ? ts_ld_relaxed(d.in_ctr) : 0u, f = d.free_ctr ? ts_ld_relaxed(d.free_ctr) : 0u;
        if ((!d.in_ctr || a >= d.in_need) && (!d.free_ctr || f >= d.free_need)) break;
        __nanosleep(200);
      }
    }
    unsigned li = 0, lf = 0, n_in_need = 0, n_free_need = 0;
    bool n_live = false, n_has_in = false, n_has_free = false;
    if (nxt < wk.count) {
      const TsDeps nd = ts_deps(P, ST, stage_i, (long long)(nxt / tiles));
      n_live = true; n_has_in = nd.in_ctr != nullptr; n_has_free = nd.free_ctr != nullptr;
      n_in_need = nd.in_need; n_free_need = nd.free_need;
      if (nd.in_ctr) li = ts_ld_relaxed(nd.in_ctr);
      if (nd.free_ctr) lf = ts_ld_relaxed(nd.free_ctr);
    }
    __syncwarp();                                                 // item i is ready: its input may be read
    int ll = lane;
    asm volatile("" : "+r"(ll));                                  // opaque per iteration (see ts_kernels.cuh: loop-invariant hoisting)
    const cpx<T>* src = ts_src(P, st.src, tr);
    cpx<T>* dst = ts_dst(P, st.dst, tr);
    if (st.kind == TS_FIRST || st.kind == TS_LATER) {
      tsw_item_phase_any<SIGN, T>(0, ll, item, st, P.Nc, src, dst, P.tw, twRs, twLs, tile);
      // the previous item's completion signal goes out HERE: its release fence waits for stores issued a whole phase ago
      // instead of stalling the warp right behind them (ncu of the first version: 2.0 warps per issue stalled on `membar`)
      if (pending) { if (lane == 0) ts_red_release(pending); pending = nullptr; }
      __syncwarp();
      tsw_item_phase_any<SIGN, T>(1, ll, item, st, P.Nc, src, dst, P.tw, twRs, twLs, tile);
    } else if (st.kind == TS_PRE) ts_pre_item<T>(lane, 32, item, st.mode, P.in + tr * 2LL * P.Nc, dst, P.N, P.Nc, P.twr);
    else if (st.kind == TS_POST) ts_post_item<T>(lane, 32, item, st.mode, src, P.out + tr * 2LL * P.Nc, P.N, P.Nc, P.twr);
    __syncwarp();                                                 // every store of the item is issued; the tile is free again
    if (pending) { if (lane == 0) ts_red_release(pending); }      // (element-wise stages: still owed)
    pending = d.done;
    cur_ready = n_live && (!n_has_in || li >= n_in_need) && (!n_has_free || lf >= n_free_need);
  }
  if (pending && lane == 0) ts_red_release(pending);
}
#endif  // __CUDACC__

}  // namespace pf
