// ts_kernels.cuh -- LARGE complex cores (Nc = 8192 .. 2^26) as a PIPELINE of tiled Stockham passes whose intermediates
// live in the 126 MB L2 of the B200 instead of HBM.
//
// Plan: Nc = R_1 * R_2 * ... * R_P, every R_i = 16*A_i with A_i from the register DFT library (1,2,3,4,5,6,8,9,10,12,15,16).
// Pass i is one autosort (Stockham, decimation in frequency) stage of radix R_i over the whole vector:
//        y[q + s*(R*p + k)] = W_Nc^{s p k} * sum_n x[b + n*m] W_R^{n k},     b = q + s*p,  m = Nc/R,  s = R_1*...*R_{i-1}
// and a WORK ITEM is a tile of 16 neighbouring columns b = 16*tile .. 16*tile+15 with all R points of each:
//   phase 1: thread (j = column, q): loads x[b + m*(q + A*i)], i < 16 -- for every n the 16 columns are ONE 128-byte run --
//            radix-16 register FFT over i -> k_a, * W_R^{q k_a}, into the CTA's exchange tile;
//   phase 2: radix-A register DFT over q -> k_b, k = k_a + 16 k_b, * W_Nc^{s p k}, stored in 128-byte runs:
//            first pass (s = 1): y[R*b + k] -- lanes run along k_a, the tile's output is one contiguous 16*R block;
//            later passes (16 | s): y[q + s*(R*p + k)] -- lanes run along the column, p is uniform over the tile.
//   The output is in natural order after the last pass: no transposes, no bit reversal, every HBM/L2 access a full line.
//
// ONE persistent kernel runs all passes (plus an element-wise pre-/post-rotation stage for real transforms and z-domain
// layouts).  Work items are dealt out statically (item k to CTA k mod gridDim.x) in the order
//        group g:  [pass 1 of transform g] [pass 2 of transform g - L] [pass 3 of transform g - 2L] ...
// so pass i+1 of a transform is scheduled ~1.5 grid-fulls of tiles after pass i finished writing it: its input is still in
// L2.  The intermediates are RINGS of 2L+1 transforms that are overwritten in place, so their dirty lines are re-written
// in L2 before they are ever evicted: HBM sees one read of x and one write of X per transform (the two-launch tiled plan of
// round 1 moved 2x that and sat at 0.42-0.45 of the HBM roofline with both launches at ~0.85 of HBM speed).
// Dependencies (per ring slot, cumulative counters): a tile of pass i+1 waits until all tiles of pass i of its transform
// are stored (release/acquire on a global counter); a tile that writes a ring slot waits until the slot's previous
// occupant has been consumed.  Every wait is on items with a SMALLER index, the grid is sized to be co-resident, so the
// scheme cannot deadlock.  Ring data is read with ld.global.cg (L2 only): a stale L1 line of a recycled slot is impossible.
// Transforms too large for L2 (> ~16 MB) run the same code with rings of 1-3 slots; the traffic then goes to HBM.
//
// Replaces, for these sizes, the cfftf1_ps/rfftf1_ps pass sweeps + finalize/preprocess + zreorder of the reference
// (src/pffft_priv_impl.h:809-1048, :1195-1462, :1158-1193), whose passes stream the whole vector through the cache
// hierarchy once per radix-4 factor.
#pragma once
#include "butterfly.cuh"
#include "generic_kernels.cuh"
#include "cta_kernels.cuh"   // brev4, ldtab

namespace pf {

enum TsKind { TS_FIRST = 0, TS_LATER = 1, TS_PRE = 2, TS_POST = 3, TS_SMALL = 4 };
enum { kTsSmallFlag = 100 };   // plan encoding: A >= kTsSmallFlag means a closing radix-(A - 100) pass without the factor 16
enum { kTsMaxStages = 6, kTsMaxRings = 5, kTsThreads = 256, kTsChunk = 4096, kTsCounterBase = 32 };

struct TsStage {
  int kind;      // TsKind
  int A;         // FFT stages: radix R = 16*A
  int mode;      // TS_PRE: LoadMode, TS_POST: StoreMode
  int tiles;     // work items per transform
  int src, dst;  // 0 = user input, 1 = user output, 2 + r = ring r
  int m;         // FFT stages: Nc / R
  int s;         // FFT stages: product of the earlier radices
  int tw_off;    // FFT stages: offset of W_R^{q k_a} (R entries, [k_a*A + q]) inside twR
};

template <typename T> struct TsParams {
  const T* in;                      // dense batch, 2*Nc scalars per transform (complex: Nc pairs; real: N = 2*Nc samples)
  T* out;
  cpx<T>* ring[kTsMaxRings];        // ring_slots * Nc complex words each
  const cpx<T>* tw;                 // exp(-2 pi i k / Nc), k < Nc
  const cpx<T>* twr;                // exp(-2 pi i k / N),  k < N/2  (real transforms)
  const cpx<T>* twR;                // per-radix tables
  unsigned* counters;               // [0] ticket; [kTsCounterBase + stage*ring_slots + slot] tiles done (cumulative)
  long long batch;
  int N, Nc;
  int nstages, lag, ring_slots, group_items;
  unsigned total_items;
  int twR_entries;                  // total entries of twR (copied to shared memory by the kernel)
  TsStage st[kTsMaxStages];
};

// read that may observe data written by another SM during this launch: L2 only
template <typename T> PF_HD cpx<T> ld_l2(const cpx<T>* p) {
#ifdef __CUDA_ARCH__
  if constexpr (sizeof(T) == 4) { const float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return mk<T>(v.x, v.y); }
  else { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return mk<T>(v.x, v.y); }
#else
  return *p;
#endif
}

// Radices with a small second factor A would leave most of the CTA idle in phase 1 (16*A threads hold a tile), so a work
// item is G = 16/A neighbouring 16-column tiles (A <= 8; one tile for the larger A): 240-256 busy threads for every radix.
template <int A> struct TsShape {
  static constexpr int R = 16 * A;
  static constexpr int G = A >= 9 ? 1 : 16 / A;              // tiles per work item
  static constexpr int COLS = 16 * G;                        // columns per work item
};
PF_HD constexpr int ts_cols_for(int A) { return A >= 9 ? 16 : 16 * (16 / A); }

// exchange-tile index of (k_a, q, column j): first pass -> phase 2 runs its lanes along k_a (XOR keeps both phases
// conflict free); later passes -> both phases run their lanes along j
template <int A, bool FIRST> PF_HD int ts_tile_idx(int ka, int q, int j) {
  return FIRST ? ((q + A * j) * 16 + (ka ^ j)) : ((ka * A + q) * 16 + j);
}

// staging buffer of a work item (input prefetch, see the kernel): entry of (tile grp, sub-sequence q, radix-16 digit i,
// column j) -- rows of 16 columns = 128 bytes, the 16 rows of one (grp, q) contiguous
template <int A> PF_HD int ts_stage_idx(int grp, int q, int i, int j) { return ((grp * A + q) * 16 + i) * 16 + j; }
// the r-th (r < 8) 16-byte piece thread t copies for the item: the 16 lanes that share (grp, q) copy exactly the 16 rows x
// 128 bytes the same 16 lanes consume in phase 1, so a __syncwarp is all that separates the copy from its use.
// Returns false when the thread has nothing to copy; *g = source element (transform-relative), *d = staging entry.
template <int A> PF_HD bool ts_stage_piece(int t, int r, int b0, int m, long long* g, int* d) {
  using S = TsShape<A>;
  const int j = t & 15, qq = t >> 4, q = qq % A, grp = qq / A;
  const int bg = b0 + 16 * grp;
  if (grp >= S::G || bg >= m) return false;
  const int i = (j >> 3) + 2 * r, c2 = 2 * (j & 7);
  *g = (long long)bg + c2 + (long long)m * (q + A * i);
  *d = ts_stage_idx<A>(grp, q, i, c2);
  return true;
}

// ---- phase 1 (both kinds): thread t -> column j = t & 15 of tile grp = (t >> 4) / A, sub-sequence q = (t >> 4) % A
// STAGED: the item's input already sits in the staging buffer `src` (ts_stage_idx layout) instead of global memory
template <int A, bool FIRST, int SIGN, typename T, bool STAGED = false>
PF_HD void ts_phase1(int t, int b0, const cpx<T>* src /* transform base, or the staging buffer */, int m, const cpx<T>* twR, cpx<T>* tile) {
  using S = TsShape<A>;
  const int j = t & 15, qq = t >> 4, q = qq % A, grp = qq / A;
  const int bg = b0 + 16 * grp;
  if (grp >= S::G || bg >= m) return;
  cpx<T> v[16];
  if (STAGED) {
#pragma unroll
    for (int p = 0; p < 16; ++p) v[p] = src[ts_stage_idx<A>(grp, q, brev4(p), j)];
  } else {
    const cpx<T>* c = src + bg + j + (long long)m * q;
#pragma unroll
    for (int p = 0; p < 16; ++p) v[p] = ld_l2(c + (long long)m * (A * brev4(p)));
  }
  reg_fft<16, SIGN>(v);
  cpx<T>* tl = tile + grp * (16 * S::R);
  tl[ts_tile_idx<A, FIRST>(0, q, j)] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tl[ts_tile_idx<A, FIRST>(ka, q, j)] = cmul_dir<SIGN>(v[ka], twR[ka * A + q]);   // (shared-memory copy in the kernel)
}
// ---- phase 2, first pass (s = 1, p = b): y[R*b + k] = W_Nc^{b k} (...),  W^{b k} = W^{b k_a} * W^{16 b k_b}
template <int A, int SIGN, typename T>
PF_HD void ts_phase2_first(int t, int b0, int m, const cpx<T>* tw, const cpx<T>* tile, cpx<T>* dst /* transform base */) {
  using S = TsShape<A>;
  const int ka = t & 15, j = t >> 4;
#pragma unroll 1
  for (int grp = 0; grp < S::G; ++grp) {
    const int b = b0 + 16 * grp + j;
    if (b0 + 16 * grp >= m) break;
    const cpx<T>* tl = tile + grp * (16 * S::R);
    cpx<T> u[A];
#pragma unroll
    for (int q = 0; q < A; ++q) u[q] = tl[ts_tile_idx<A, true>(ka, q, j)];
    dft_small<A, SIGN>(u);
    const cpx<T> w1 = ldtab(tw + b * ka);
    cpx<T>* o = dst + (long long)S::R * b + ka;
    o[0] = cmul_dir<SIGN>(u[0], w1);
#pragma unroll
    for (int kb = 1; kb < A; ++kb) o[16 * kb] = cmul_dir<SIGN>(u[kb], cmul(w1, ldtab(tw + 16 * b * kb)));
  }
}
// ---- phase 2, later passes (16 | s): p and the twiddle exponent base are uniform over a tile
template <int A, int SIGN, typename T>
PF_HD void ts_phase2_later(int t, int b0, int m, int s, const cpx<T>* tw, const cpx<T>* tile, cpx<T>* dst) {
  using S = TsShape<A>;
  const int j = t & 15, ka = t >> 4;
#pragma unroll 1
  for (int grp = 0; grp < S::G; ++grp) {
    const int bg = b0 + 16 * grp;
    if (bg >= m) break;
    const cpx<T>* tl = tile + grp * (16 * S::R);
    cpx<T> u[A];
#pragma unroll
    for (int q = 0; q < A; ++q) u[q] = tl[ts_tile_idx<A, false>(ka, q, j)];
    dft_small<A, SIGN>(u);
    const int pq = bg / s, e0 = pq * s;                     // e0 = s*p
    cpx<T>* o = dst + (bg - e0 + j) + (long long)s * ((long long)S::R * pq + ka);
    const long long ks = 16LL * s;
    if (e0 == 0) {
#pragma unroll
      for (int kb = 0; kb < A; ++kb) o[ks * kb] = u[kb];
    } else {
#pragma unroll
      for (int kb = 0; kb < A; ++kb) o[ks * kb] = cmul_dir<SIGN>(u[kb], ldtab(tw + (long long)e0 * (ka + 16 * kb)));
    }
  }
}

// one FFT work item, phase by phase (the kernel puts a CTA barrier between them; tests/emu steps them lane by lane)
template <int A, bool FIRST, int SIGN, typename T>
PF_HD void ts_item_phase(int phase, int t, int item, const TsStage& st, const cpx<T>* src, cpx<T>* dst,
                         const cpx<T>* tw, const cpx<T>* twR, cpx<T>* tile) {
  const int b0 = TsShape<A>::COLS * item;
  if (phase == 0) ts_phase1<A, FIRST, SIGN, T>(t, b0, src, st.m, twR + st.tw_off, tile);
  else if (phase == 2) ts_phase1<A, FIRST, SIGN, T, true>(t, b0, src, st.m, twR + st.tw_off, tile);   // src = staging buffer
  else if (FIRST) ts_phase2_first<A, SIGN, T>(t, b0, st.m, tw, tile, dst);
  else ts_phase2_later<A, SIGN, T>(t, b0, st.m, st.s, tw, tile, dst);
}
template <bool FIRST, int SIGN, typename T>
PF_HD void ts_item_phase_any(int phase, int t, int item, const TsStage& st, const cpx<T>* src, cpx<T>* dst,
                             const cpx<T>* tw, const cpx<T>* twR, cpx<T>* tile) {
  switch (st.A) {
#define PF_TS(a) case a: ts_item_phase<a, FIRST, SIGN, T>(phase, t, item, st, src, dst, tw, twR, tile); break;
    PF_TS(1) PF_TS(2) PF_TS(3) PF_TS(4) PF_TS(5) PF_TS(6) PF_TS(8) PF_TS(9) PF_TS(10) PF_TS(12) PF_TS(15) PF_TS(16)
#undef PF_TS
    default: break;
  }
}

// ---- closing pass of a small radix A (no factor 16; cores with fewer than 4 factors of two per pass, e.g. 384000 =
// 240 x 160 x 10): one column per thread, no exchange.  Work item = 256 neighbouring columns.
template <int A, int SIGN, typename T>
PF_HD void ts_small_item(int t, int item, int m, int s, const cpx<T>* src, cpx<T>* dst, const cpx<T>* tw) {
  const int b = 256 * item + t;
  if (b >= m) return;
  cpx<T> u[A];
#pragma unroll
  for (int n = 0; n < A; ++n) u[n] = ld_l2(src + b + (long long)m * n);
  dft_small<A, SIGN>(u);
  const int pq = b / s, e0 = pq * s;
  cpx<T>* o = dst + (b - e0) + (long long)s * ((long long)A * pq);
  o[0] = u[0];
#pragma unroll
  for (int k = 1; k < A; ++k) o[(long long)s * k] = e0 ? cmul_dir<SIGN>(u[k], ldtab(tw + (long long)e0 * k)) : u[k];
}
template <int SIGN, typename T>
PF_HD void ts_small_item_any(int t, int item, const TsStage& st, const cpx<T>* src, cpx<T>* dst, const cpx<T>* tw) {
  switch (st.A) {
#define PF_TS(a) case a: ts_small_item<a, SIGN, T>(t, item, st.m, st.s, src, dst, tw); break;
    PF_TS(2) PF_TS(3) PF_TS(4) PF_TS(5) PF_TS(6) PF_TS(8) PF_TS(9) PF_TS(10) PF_TS(12) PF_TS(15)
#undef PF_TS
    default: break;
  }
}

// element-wise stages: chunk c of kTsChunk core elements of one transform
template <typename T>
PF_HD void ts_pre_item(int t, int nthreads, int chunk, int mode, const T* ibase, cpx<T>* core, int N, int Nc, const cpx<T>* twr) {
  const int hi = (chunk + 1) * kTsChunk < Nc ? (chunk + 1) * kTsChunk : Nc;
  for (int i = chunk * kTsChunk + t; i < hi; i += nthreads) {
    cpx<T> v;
    if (mode == L_C_Z) v = load_core<L_C_Z, T>(ibase, i, N, Nc, twr, -1, true);
    else if (mode == L_R_ORD) v = load_core<L_R_ORD, T>(ibase, i, N, Nc, twr, -1, true);
    else v = load_core<L_R_Z, T>(ibase, i, N, Nc, twr, -1, true);
    core[i] = v;
  }
}
template <typename T>
PF_HD void ts_post_item(int t, int nthreads, int chunk, int mode, const cpx<T>* core, T* obase, int N, int Nc, const cpx<T>* twr) {
  const int hi = (chunk + 1) * kTsChunk < Nc ? (chunk + 1) * kTsChunk : Nc;
  for (int k = chunk * kTsChunk + t; k < hi; k += nthreads) {
    if (mode == S_C_Z) store_core<S_C_Z, T, true>(obase, core, k, N, Nc, twr, N, true);
    else if (mode == S_R_ORD) store_core<S_R_ORD, T, true>(obase, core, k, N, Nc, twr, N, true);
    else store_core<S_R_Z, T, true>(obase, core, k, N, Nc, twr, N, true);
  }
}

// ticket -> (stage, transform, tile); false when the slot is padding (pipeline fill / drain)
// ST: the stage table (the kernel works on a shared-memory copy: dynamic indexing into the by-value parameter struct would
// make the compiler copy it to local memory)
template <typename T>
PF_HD bool ts_decode(const TsParams<T>& P, const TsStage* ST, unsigned ticket, int* stage, long long* tr, int* item) {
  const unsigned g = ticket / (unsigned)P.group_items;
  int r = (int)(ticket - g * (unsigned)P.group_items);
  int i = 0;
  while (i < P.nstages - 1 && r >= ST[i].tiles) { r -= ST[i].tiles; ++i; }
  const long long t = (long long)g - (long long)i * P.lag;
  *stage = i; *tr = t; *item = r;
  return t >= 0 && t < P.batch;
}
// ring slot of transform tr: 32-bit arithmetic (tr < 2^31, see ts_run) -- a 64-bit modulo is a ~100-instruction subroutine,
// and the first versions of the kernel ran five of them per thread and work item (40 % of all executed instructions, with
// thread 0's copies on the critical path of every barrier)
// STAGE-SPECIALISED WORK DISTRIBUTION (round 2b; used by k_tsw_pipeline -- measured on k_ts_pipeline too and taken out again,
// see the note at the end of this comment).  Worker g of W (a warp of k_tsw_pipeline) serves
// ONE stage, s = g mod nstages, as the (g div nstages)-th of the n_s workers there, and takes that stage's work items
// q = j, j + n_s, ... in order (q = transform * tiles + tile).  The first versions dealt all stages out to all workers in one
// interleaved order ("pass 2 of transform g - L behind pass 1 of transform g"): a worker whose pass-2 item had to wait for
// its input could not start its next pass-1 item, whose consumers then waited in turn -- head-of-line blocking; ncu counted
// 36 polls per work item in the CTA kernel and 139 in the warp kernel.  Now the producers of a stage never wait for its
// consumers (only for a free ring slot, 2L+1 transforms ahead), every worker waits only for work of an EARLIER transform or
// stage, and all workers are co-resident: no deadlock.
// On the CTA kernel the polls disappeared and the time did not change (the waits had been hidden behind the other CTAs of
// the SM): 16384 / 65536 / 2^17 / 2^20 / 2^24 / 2^26: 0.45 / 0.40 / 0.28 / 0.25 / 0.13 / 0.07 of the roofline against 0.45 / 0.44 /
// 0.26 / 0.27 / 0.23 / 0.16 interleaved -- a stage's producers need ring slots for every transform they hold in flight, which
// transforms of 128 MB do not have.  k_ts_pipeline went back to the interleaved order (profiles/r02b_large_n.md).
struct TsWorker { int stage; unsigned q0, step, count; };
template <typename T> PF_HD TsWorker ts_worker(const TsParams<T>& P, const TsStage* ST, unsigned g, unsigned W) {
  TsWorker w;
  const unsigned ns = (unsigned)P.nstages;
  w.stage = (int)(g % ns);
  w.q0 = g / ns;
  w.step = (W - (unsigned)w.stage + ns - 1) / ns;
  w.count = (unsigned)P.batch * (unsigned)ST[w.stage].tiles;      // < 2^32 - W (ts_run)
  return w;
}

template <typename T> PF_HD unsigned ts_slot(const TsParams<T>& P, long long tr) { return (unsigned)tr % (unsigned)P.ring_slots; }
template <typename T> PF_HD const cpx<T>* ts_src(const TsParams<T>& P, int which, long long tr) {
  if (which == 0) return reinterpret_cast<const cpx<T>*>(P.in) + tr * P.Nc;
  return P.ring[which - 2] + (long long)ts_slot(P, tr) * P.Nc;
}
template <typename T> PF_HD cpx<T>* ts_dst(const TsParams<T>& P, int which, long long tr) {
  if (which == 1) return reinterpret_cast<cpx<T>*>(P.out) + tr * P.Nc;
  return P.ring[which - 2] + (long long)ts_slot(P, tr) * P.Nc;
}

#ifdef __CUDACC__
PF_D unsigned ts_ld_relaxed(const unsigned* p) {                    // L2 read without the L1 invalidation of an acquire
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PF_D void ts_red_release(unsigned* p) {                             // completion signal: earlier writes of the CTA first
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// ---- input prefetch (float): the NEXT work item's 32 KB of input are copied global/L2 -> shared by 16-byte cp.async.cg
// (LDGSTS, L2 only like ld.cg, no registers held) while the current item runs its phase 2
PF_D void ts_cp_async_cg16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
PF_D void ts_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int A, typename T> PF_D void ts_prefetch_item(int t, int item, int m, const cpx<T>* src, cpx<T>* stage) {
  const int b0 = TsShape<A>::COLS * item;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    long long g; int d;
    if (ts_stage_piece<A>(t, r, b0, m, &g, &d)) ts_cp_async_cg16(stage + d, src + g);
  }
}
template <typename T> PF_D void ts_prefetch_any(int t, int item, const TsStage& st, const cpx<T>* src, cpx<T>* stage) {
  switch (st.A) {
#define PF_TS(a) case a: ts_prefetch_item<a, T>(t, item, st.m, src, stage); break;
    PF_TS(1) PF_TS(2) PF_TS(3) PF_TS(4) PF_TS(5) PF_TS(6) PF_TS(8) PF_TS(9) PF_TS(10) PF_TS(12) PF_TS(15) PF_TS(16)
#undef PF_TS
    default: break;
  }
}

// dependency counters of a live work item: `in_need` tiles of the producing stage, `free_need` tiles of the consuming
// stage's previous occupant of the ring slot (nullptr pointers: no such dependency)
struct TsDeps { const unsigned* in_ctr; unsigned in_need; const unsigned* free_ctr; unsigned free_need; unsigned* done; };
template <typename T> PF_D TsDeps ts_deps(const TsParams<T>& P, const TsStage* ST, int stage, long long tr) {
  const unsigned slot = ts_slot(P, tr);
  const unsigned gen = (unsigned)tr / (unsigned)P.ring_slots;
  unsigned* base = P.counters + kTsCounterBase + slot;
  TsDeps d;
  d.done = base + stage * P.ring_slots;
  d.in_ctr = stage > 0 ? base + (stage - 1) * P.ring_slots : nullptr;
  d.in_need = stage > 0 ? (gen + 1u) * (unsigned)ST[stage - 1].tiles : 0u;
  const bool fr = stage + 1 < P.nstages && gen > 0;
  d.free_ctr = fr ? base + (stage + 1) * P.ring_slots : nullptr;
  d.free_need = fr ? gen * (unsigned)ST[stage + 1].tiles : 0u;
  return d;
}

// THE PERSISTENT LOOP.  What the hardware profiles of the first versions showed (profiles/r02_large_n.md) and what is
// left of five variants:
//   * work items are dealt out STATICALLY, item k to CTA k mod gridDim.x.  (Atomic tickets, one or two kept in hand to hide
//     the atomic's latency, widen the window of unfinished items by a whole grid per ticket in hand; the pipeline lag no
//     longer covered it and 45 % of the items polled.)  The deadlock argument is unchanged: a CTA only ever waits for
//     items with a smaller index, and the grid is co-resident;
//   * the counters of item i+1 are read (ld.relaxed, both at once) at the top of item i and looked at when item i's first
//     phase is done -- by then they have long arrived and, with the pipeline lag, they are satisfied: item i+1 starts
//     without waiting for anything.  Only if the early look failed does its top poll (still relaxed);
//   * nothing invalidates L1 (ld.acquire / __threadfence compile to CCTL.IVALL, which threw the twiddle tables of all
//     resident CTAs out once per item): ring data is only ever read with ld.cg (L2), tables are immutable, the per-radix
//     tables live in shared memory, the completion signal is one red.release issued by a thread of another warp than the
//     one that handles the counters, so neither waits for the other;
//   * the thread index is made opaque once per iteration, see below (0 spill bytes instead of 600-780).
// Measured and dropped (slower): cp.async prefetch of the next item by all threads (0.25-0.30 of the roofline against
// 0.35-0.40), a warp-specialised producer staging items with 1-D TMA bulk copies (0.14: 128-byte cp.async.bulk cost
// 10-20 ns each) or with cp.async (0.19), a producer warp for the control path only (0.26-0.35).
//   * PRE (round 2b, float): the input of item i+1 is prefetched into a second shared buffer while item i runs its phase 2
//     (ts_prefetch_item), issued once thread 0's early look has shown that item i+1's dependencies are met.  An item's
//     timeline was [wait for 16 loads] [radix 16] [barrier] [twiddle loads, radix A, stores] with three CTAs per SM to
//     overlap it: 30 % issue utilisation.  With the prefetch phase 1 starts from shared memory.
template <typename T, int SIGN, int MINB, bool PRE>
__global__ void __launch_bounds__(kTsThreads, MINB) k_ts_pipeline(const __grid_constant__ TsParams<T> P) {
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<T>* tile = reinterpret_cast<cpx<T>*>(pf_smem_raw);
  cpx<T>* stage = tile + 16 * 256;                              // PRE only
  cpx<T>* twRs = PRE ? stage + 16 * 256 : stage;                // per-radix tables of every pass (<= 4 x 256 entries)
  __shared__ int s_cur_ready, s_next_ready;
  __shared__ TsStage ST[kTsMaxStages];                          // (dynamic indexing into the by-value parameter would
  const int t = threadIdx.x;                                    //  make the compiler copy it to local memory)
  for (int i = t; i < P.twR_entries; i += kTsThreads) twRs[i] = P.twR[i];
  if (t == 0) {
#pragma unroll
    for (int i = 0; i < kTsMaxStages; ++i) ST[i] = P.st[i];       // constant indices: plain constant-bank reads
    s_cur_ready = 0;
  }
  __syncthreads();
  bool staged = false;                                            // CTA-uniform: the current item's input is in `stage`
  for (unsigned cur = blockIdx.x; cur < P.total_items; cur += gridDim.x) {
    const unsigned nxt = cur + gridDim.x;                         // (total_items + gridDim.x < 2^32: checked by the host)
    // ---- readiness of item i: known from the early look of the previous iteration (s_cur_ready, CTA-uniform after the
    // closing barrier) in all but a few per cent of the items -- only then does thread 0 poll, behind a barrier of its own
    if (!s_cur_ready) {
      if (t == 0) {
        int stage_p, item_p; long long tr_p;
        if (ts_decode(P, ST, cur, &stage_p, &tr_p, &item_p)) {
          const TsDeps d = ts_deps(P, ST, stage_p, tr_p);
          for (;;) {
            const unsigned a = d.in_ctr ? ts_ld_relaxed(d.in_ctr) : 0u, f = d.free_ctr ? ts_ld_relaxed(d.free_ctr) : 0u;
            if ((!d.in_ctr || a >= d.in_need) && (!d.free_ctr || f >= d.free_need)) break;
            __nanosleep(100);
          }
        }
      }
      __syncthreads();
    }
    // ---- thread 0: early look at item i+1 -- its counters are read now (off everybody else's critical path) and looked at
    // after phase 1
    unsigned li = 0, lf = 0, n_in_need = 0, n_free_need = 0;
    bool n_live = false, n_has_in = false, n_has_free = false;
    if (t == 0) {
      int nstage, nitem; long long ntr;
      if (nxt < P.total_items && ts_decode(P, ST, nxt, &nstage, &ntr, &nitem)) {
        const TsDeps nd = ts_deps(P, ST, nstage, ntr);
        n_live = true; n_has_in = nd.in_ctr != nullptr; n_has_free = nd.free_ctr != nullptr;
        n_in_need = nd.in_need; n_free_need = nd.free_need;
        if (nd.in_ctr) li = ts_ld_relaxed(nd.in_ctr);             // in flight during phase 1
        if (nd.free_ctr) lf = ts_ld_relaxed(nd.free_ctr);
      }
    }
    int stage_i, item; long long tr;
    const bool live = ts_decode(P, ST, cur, &stage_i, &tr, &item);
    const TsStage& st = ST[stage_i];
    const bool fft = live && (st.kind == TS_FIRST || st.kind == TS_LATER);
    // the thread index is made opaque per iteration: otherwise the compiler hoists the per-thread index arithmetic of ALL
    // 24 radix bodies (t % A, t / A, tile offsets ...) out of the persistent loop and spills ~150 values to local memory
    int tt = t;
    asm volatile("" : "+r"(tt));
    if (fft) {
      const cpx<T>* src = ts_src(P, st.src, tr);
      int ph = 0;
      if constexpr (PRE) if (staged) { ts_cp_async_wait_all(); __syncwarp(); src = stage; ph = 2; }   // copied by the same 16 lanes that read it
      if (st.kind == TS_FIRST) ts_item_phase_any<true, SIGN, T>(ph, tt, item, st, src, (cpx<T>*)nullptr, P.tw, twRs, tile);
      else ts_item_phase_any<false, SIGN, T>(ph, tt, item, st, src, (cpx<T>*)nullptr, P.tw, twRs, tile);
    }
    if (t == 0)                                                   // is item i+1 known to be ready?
      s_next_ready = n_live && (!n_has_in || li >= n_in_need) && (!n_has_free || lf >= n_free_need);
    __syncthreads();                                              // tile complete; nobody reads `stage` any more
    bool pre_next = false;
    if constexpr (PRE) if (s_next_ready) {                        // item i+1 is live and its input complete: start fetching it
      int nstage, nitem; long long ntr;
      ts_decode(P, ST, nxt, &nstage, &ntr, &nitem);
      const TsStage& nst = ST[nstage];
      if (nst.kind == TS_FIRST || nst.kind == TS_LATER) {
        ts_prefetch_any<T>(tt, nitem, nst, ts_src(P, nst.src, ntr), stage);
        pre_next = true;
      }
    }
    if (live) {
      const cpx<T>* src = ts_src(P, st.src, tr);
      cpx<T>* dst = ts_dst(P, st.dst, tr);
      if (st.kind == TS_FIRST) ts_item_phase_any<true, SIGN, T>(1, tt, item, st, src, dst, P.tw, twRs, tile);
      else if (st.kind == TS_LATER) ts_item_phase_any<false, SIGN, T>(1, tt, item, st, src, dst, P.tw, twRs, tile);
      else if (st.kind == TS_SMALL) ts_small_item_any<SIGN, T>(tt, item, st, src, dst, P.tw);
      else if (st.kind == TS_PRE) ts_pre_item<T>(t, kTsThreads, item, st.mode, P.in + tr * 2LL * P.Nc, dst, P.N, P.Nc, P.twr);
      else ts_post_item<T>(t, kTsThreads, item, st.mode, src, P.out + tr * 2LL * P.Nc, P.N, P.Nc, P.twr);
    }
    if (t == 0) s_cur_ready = s_next_ready;
    staged = pre_next;
    __syncthreads();                                              // every store of the item is issued; tile is free again
    if (t == kTsThreads - 32 && live) ts_red_release(ts_deps(P, ST, stage_i, tr).done);   // another warp than thread 0's
  }
}
#endif  // __CUDACC__

}  // namespace pf
