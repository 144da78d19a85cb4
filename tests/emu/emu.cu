// emu.cu -- DEVELOPMENT HARNESS, not part of the product.
// Steps the PF_HD device functions of pffft_b200/csrc lane-by-lane on the CPU so index math and
// butterfly algebra can be checked in the GPU-less build container before a gpurun call.
// Nothing in pffft_b200/ links or loads this; only tests/test_emu_*.py does.
#include <vector>
#include <cstring>
#include "../../pffft_b200/csrc/generic_kernels.cuh"
#include "../../pffft_b200/csrc/plan.h"
#include "emu_fast.cuh"

using namespace pf;

template <typename T, int LM, int SM, int SIGN>
static void run_generic(XformParams<T> p) {
  std::vector<cpx<T>> A(p.Nc), B(p.Nc);
  for (long long t = 0; t < p.batch; ++t) {
    const T* ibase = p.in + t * p.in_stride;
    const long long avail = p.in_limit < 0 ? -1 : p.in_limit - t * p.in_stride;
    for (int i = 0; i < p.Nc; ++i) A[i] = load_core<LM, T>(ibase, i, p.N, p.Nc, p.twr, avail, vec_aligned<T>(ibase));
    cpx<T>* src = A.data(); cpx<T>* dst = B.data();
    int s = 1;
    for (int f = 0; f < p.nfac; ++f) {
      const int r = p.fac[f], m = p.Nc / r;
      for (int li = 0; li < 4; ++li) stockham_stage<SIGN, T>(r, src, dst, li, 4, p.Nc, s, p.magic[f], p.tw);   // 4 emulated lanes
      std::swap(src, dst); s *= r;
    }
    T* obase = p.out + t * p.out_stride;
    std::vector<cpx<T>> Z(src, src + p.Nc);   // store may alias input
    for (int k = 0; k < p.Nc; ++k) store_core<SM, T>(obase, Z.data(), k, p.N, p.Nc, p.twr, p.out_count, vec_aligned<T>(obase));
  }
}

template <typename T, int LM, int SM>
static void run_dir(XformParams<T> p, int dir) { if (dir == 0) run_generic<T, LM, SM, -1>(p); else run_generic<T, LM, SM, +1>(p); }

template <typename T>
static int emu_generic_t(int N, int transform, int dir, int lm, int sm, const T* in, T* out, long long batch,
                         long long in_stride, long long out_stride, long long in_limit, int out_count) {
  if (!pfplan::setup_size_ok(N, transform)) return -1;
  XformParams<T> p{};
  p.N = N; p.Nc = transform == 0 ? N / 2 : N;
  auto f = pfplan::factorize(p.Nc);
  p.nfac = (int)f.size(); { int prod = 1; for (int i = 0; i < p.nfac; ++i) { p.fac[i] = f[i]; p.magic[i] = stage_magic(prod); prod *= f[i]; } }
  std::vector<T> tw(2 * (size_t)p.Nc), twr(2 * (size_t)(N / 2));
  pfplan::fill_roots<T>(tw.data(), p.Nc, p.Nc);
  pfplan::fill_roots<T>(twr.data(), N / 2, N);
  p.tw = reinterpret_cast<const cpx<T>*>(tw.data()); p.twr = reinterpret_cast<const cpx<T>*>(twr.data());
  p.in = in; p.out = out; p.batch = batch; p.in_stride = in_stride; p.out_stride = out_stride;
  p.in_limit = in_limit; p.out_count = out_count;
#define CASE(L, S) if (lm == L && sm == S) { run_dir<T, L, S>(p, dir); return 0; }
  CASE(L_C_ORD, S_C_ORD) CASE(L_C_ORD, S_C_Z) CASE(L_C_Z, S_C_ORD) CASE(L_C_Z, S_C_Z)
  CASE(L_R_TIME, S_R_ORD) CASE(L_R_TIME, S_R_Z) CASE(L_R_ORD, S_R_TIME) CASE(L_R_Z, S_R_TIME)
#undef CASE
  return -2;
}

extern "C" int emu_generic(int prec, int N, int transform, int dir, int lm, int sm, const void* in, void* out,
                           long long batch, long long in_stride, long long out_stride, long long in_limit, int out_count) {
  if (prec == 0) return emu_generic_t<float>(N, transform, dir, lm, sm, (const float*)in, (float*)out, batch, in_stride, out_stride, in_limit, out_count);
  return emu_generic_t<double>(N, transform, dir, lm, sm, (const double*)in, (double*)out, batch, in_stride, out_stride, in_limit, out_count);
}

// z-domain index maps, for a direct comparison with the reference's pffft_zreorder
extern "C" int emu_zpos(int real, int k, int N) { return real ? zpos_real(k, N) : zpos_complex(k, N); }

// register FFT check: x[N] natural order in, X[N] natural order out
template <int N, int SIGN> static void regfft_run(const float* in, float* out) {
  cpx<float> v[N];
  constexpr int bits = ct::ilog2(N);
  for (int p = 0; p < N; ++p) { int n = ct::bitrev(p, bits); v[p] = mk<float>(in[2 * n], in[2 * n + 1]); }
  reg_fft<N, SIGN>(v);
  for (int k = 0; k < N; ++k) { out[2 * k] = v[k].x; out[2 * k + 1] = v[k].y; }
}
extern "C" int emu_regfft(int N, int dir, const float* in, float* out) {
#define RF(n) if (N == n) { if (dir == 0) regfft_run<n, -1>(in, out); else regfft_run<n, +1>(in, out); return 0; }
  RF(2) RF(4) RF(8) RF(16) RF(32) RF(64)
#undef RF
  return -1;
}

// overlap-save block algebra (host logic of pffastconv_apply) for a CPU check against the reference
extern "C" long long emu_fastconv_produced(long long inputLen, int Nfft, int filterLen, int flush, int even_out) {
  return pfplan::plan_blocks(inputLen, Nfft, filterLen, flush, even_out != 0).produced;
}

// ---- tiled Stockham pipeline (ts_kernels.cuh): the persistent kernel's ticket loop stepped on the CPU.  `window` tickets
// are "in flight" at a time (as many as the GPU has resident CTAs); the next one to run is picked AT RANDOM among those whose
// dependency counters are satisfied, so the emulation checks that (1) the counters alone are sufficient for correct data
// under any interleaving the hardware may produce and (2) some in-flight ticket is always runnable (no deadlock).
#include "../../pffft_b200/csrc/ts_plan.h"
// the kernel's input prefetch (float): every thread's eight 16-byte cp.async pieces, as plain copies into a poisoned buffer
template <int A, typename T> static void ts_emu_stage_a(int item, int m, const cpx<T>* src, cpx<T>* stage) {
  for (int t = 0; t < kTsThreads; ++t)
    for (int r = 0; r < 8; ++r) {
      long long g; int d;
      if (ts_stage_piece<A>(t, r, TsShape<A>::COLS * item, m, &g, &d)) { stage[d] = src[g]; stage[d + 1] = src[g + 1]; }
    }
}
template <typename T> static void ts_emu_stage(int item, const TsStage& st, const cpx<T>* src, cpx<T>* stage) {
  switch (st.A) {
#define PF_TS(a) case a: ts_emu_stage_a<a, T>(item, st.m, src, stage); break;
    PF_TS(1) PF_TS(2) PF_TS(3) PF_TS(4) PF_TS(5) PF_TS(6) PF_TS(8) PF_TS(9) PF_TS(10) PF_TS(12) PF_TS(15) PF_TS(16)
#undef PF_TS
    default: break;
  }
}
// `window` workers (ts_worker: each serves one stage, its work items in order); the next item to run is picked AT RANDOM among
// the workers whose current item has its dependency counters satisfied.  run(stage, tr, item) executes one work item.
template <typename T, typename Run>
static int ts_schedule(TsParams<T>& P, int window, unsigned seed, Run run) {
  std::vector<unsigned> counters(kTsCounterBase + (size_t)kTsMaxStages * P.ring_slots, 0u);
  if (window < P.nstages) window = P.nstages;                     // every stage needs a worker (ts_run enforces the same)
  std::vector<TsWorker> wk(window);
  std::vector<unsigned> cur(window);
  for (int g = 0; g < window; ++g) { wk[g] = ts_worker(P, P.st, (unsigned)g, (unsigned)window); cur[g] = wk[g].q0; }
  uint32_t rs = seed * 2654435761u + 12345u;
  auto ready = [&](int stage, long long tr) {
    const int slot = (int)(tr % P.ring_slots); const unsigned gen = (unsigned)(tr / P.ring_slots);
    const unsigned* base = counters.data() + kTsCounterBase + slot;
    if (stage > 0 && base[(stage - 1) * P.ring_slots] < (gen + 1u) * (unsigned)P.st[stage - 1].tiles) return false;
    if (stage + 1 < P.nstages && gen > 0 && base[(stage + 1) * P.ring_slots] < gen * (unsigned)P.st[stage + 1].tiles) return false;
    return true;
  };
  for (;;) {
    std::vector<int> cand; bool any = false;
    for (int g = 0; g < window; ++g) {
      if (cur[g] >= wk[g].count) continue;
      any = true;
      if (ready(wk[g].stage, (long long)(cur[g] / (unsigned)P.st[wk[g].stage].tiles))) cand.push_back(g);
    }
    if (!any) break;
    if (cand.empty()) return -10;                                  // deadlock
    rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
    const int g = cand[rs % cand.size()];
    const int stage = wk[g].stage;
    const unsigned tiles = (unsigned)P.st[stage].tiles;
    const long long tr = (long long)(cur[g] / tiles);
    const int item = (int)(cur[g] - (unsigned)tr * tiles);
    cur[g] += wk[g].step;
    run(stage, tr, item, rs);
    counters[kTsCounterBase + stage * P.ring_slots + (int)(tr % P.ring_slots)] += 1;
  }
  return 0;
}
template <typename T, int SIGN>
static int ts_emulate(TsParams<T>& P, int window, unsigned seed) {
  std::vector<cpx<T>> stagebuf(16 * 256);
  std::vector<unsigned> counters(kTsCounterBase + (size_t)kTsMaxStages * P.ring_slots, 0u);
  std::vector<cpx<T>> tile(16 * 256);
  std::vector<unsigned> flight;
  unsigned next = 0;
  uint32_t rs = seed * 2654435761u + 12345u;
  auto ready = [&](unsigned ticket) {
    int stage, item; long long tr;
    if (!ts_decode(P, P.st, ticket, &stage, &tr, &item)) return true;
    const int slot = (int)(tr % P.ring_slots); const unsigned gen = (unsigned)(tr / P.ring_slots);
    const unsigned* base = counters.data() + kTsCounterBase + slot;
    if (stage > 0 && base[(stage - 1) * P.ring_slots] < (gen + 1u) * (unsigned)P.st[stage - 1].tiles) return false;
    if (stage + 1 < P.nstages && gen > 0 && base[(stage + 1) * P.ring_slots] < gen * (unsigned)P.st[stage + 1].tiles) return false;
    return true;
  };
  while (next < P.total_items || !flight.empty()) {
    while ((int)flight.size() < window && next < P.total_items) flight.push_back(next++);
    std::vector<int> cand;
    for (int i = 0; i < (int)flight.size(); ++i) if (ready(flight[i])) cand.push_back(i);
    if (cand.empty()) return -10;                                  // deadlock
    rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
    const int pick = cand[rs % cand.size()];
    const unsigned cur = flight[pick];
    flight.erase(flight.begin() + pick);
    int stage, item; long long tr;
    if (!ts_decode(P, P.st, cur, &stage, &tr, &item)) continue;
    const TsStage& st = P.st[stage];
    const cpx<T>* src = ts_src(P, st.src, tr);
    cpx<T>* dst = ts_dst(P, st.dst, tr);
    if (st.kind == TS_FIRST || st.kind == TS_LATER) {
      // every other item (float) takes the prefetched path: phase 1 reads the staging buffer instead of global memory
      rs ^= rs << 13; rs ^= rs >> 17; rs ^= rs << 5;
      const bool staged = sizeof(T) == 4 && (rs & 1u);
      if (staged) {
        for (auto& e : stagebuf) { e.x = (T)NAN; e.y = (T)NAN; }
        ts_emu_stage<T>(item, st, src, stagebuf.data());
      }
      for (int phase = 0; phase < 2; ++phase)
        for (int t = 0; t < kTsThreads; ++t) {
          const int ph = (phase == 0 && staged) ? 2 : phase;
          const cpx<T>* s1 = (phase == 0 && staged) ? stagebuf.data() : src;
          if (st.kind == TS_FIRST) ts_item_phase_any<true, SIGN, T>(ph, t, item, st, s1, dst, P.tw, P.twR, tile.data());
          else ts_item_phase_any<false, SIGN, T>(ph, t, item, st, s1, dst, P.tw, P.twR, tile.data());
        }
    } else if (st.kind == TS_SMALL) {
      for (int t = 0; t < kTsThreads; ++t) ts_small_item_any<SIGN, T>(t, item, st, src, dst, P.tw);
    } else if (st.kind == TS_PRE) {
      for (int t = 0; t < kTsThreads; ++t) ts_pre_item<T>(t, kTsThreads, item, st.mode, P.in + tr * 2LL * P.Nc, dst, P.N, P.Nc, P.twr);
    } else {
      for (int t = 0; t < kTsThreads; ++t) ts_post_item<T>(t, kTsThreads, item, st.mode, src, P.out + tr * 2LL * P.Nc, P.N, P.Nc, P.twr);
    }
    counters[kTsCounterBase + stage * P.ring_slots + (int)(tr % P.ring_slots)] += 1;
  }
  return 0;
}
// ---- warp-sized work items (tsw_kernels.cuh): same protocol, 32 lanes per item, two phases around a __syncwarp
template <typename T, int SIGN>
static int tsw_emulate(TsParams<T>& P, const cpx<T>* twL, int window, unsigned seed) {
  std::vector<cpx<T>> tile(kTswTileMax);
  int bad = 0;
  const int rc = ts_schedule<T>(P, window, seed, [&](int stage, long long tr, int item, uint32_t) {
    const TsStage& st = P.st[stage];
    const cpx<T>* src = ts_src(P, st.src, tr);
    cpx<T>* dst = ts_dst(P, st.dst, tr);
    if (st.kind == TS_FIRST || st.kind == TS_LATER) {
      for (auto& e : tile) { e.x = (T)NAN; e.y = (T)NAN; }
      for (int phase = 0; phase < 2; ++phase)
        for (int lane = 0; lane < 32; ++lane) tsw_item_phase_any<SIGN, T>(phase, lane, item, st, P.Nc, src, dst, P.tw, P.twR, twL, tile.data());
    } else if (st.kind == TS_PRE) {
      for (int lane = 0; lane < 32; ++lane) ts_pre_item<T>(lane, 32, item, st.mode, P.in + tr * 2LL * P.Nc, dst, P.N, P.Nc, P.twr);
    } else if (st.kind == TS_POST) {
      for (int lane = 0; lane < 32; ++lane) ts_post_item<T>(lane, 32, item, st.mode, src, P.out + tr * 2LL * P.Nc, P.N, P.Nc, P.twr);
    } else bad = 1;
  });
  return bad ? -11 : rc;
}
template <typename T>
static int emu_ts_t(int N, int transform, int dir, int ordered, const T* in, T* out, long long batch, int lag, int window, unsigned seed,
                    bool wmode = false) {
  const int Nc = transform == 0 ? N / 2 : N;
  int Pn = 0, A[4], tw_off[4];
  if (!ts_factorize(Nc, &Pn, A)) return -1;
  std::vector<T> tw(2 * (size_t)Nc), twr(2 * (size_t)(N / 2));
  pfplan::fill_roots<T>(tw.data(), Nc, Nc);
  pfplan::fill_roots<T>(twr.data(), N / 2, N);
  const std::vector<T> twR = ts_radix_tables<T>(Pn, A, tw_off);
  TsParams<T> P;
  memset(&P, 0, sizeof(P));
  P.in = in; P.out = out; P.batch = batch; P.N = N; P.Nc = Nc;
  P.tw = reinterpret_cast<const cpx<T>*>(tw.data()); P.twr = reinterpret_cast<const cpx<T>*>(twr.data());
  P.twR = reinterpret_cast<const cpx<T>*>(twR.data());
  P.lag = lag; P.ring_slots = lag > 0 ? 2 * lag + 1 : 1;
  std::vector<std::vector<cpx<T>>> rings(kTsMaxRings, std::vector<cpx<T>>((size_t)P.ring_slots * Nc));
  for (int i = 0; i < kTsMaxRings; ++i) P.ring[i] = rings[i].data();
  int lm, sm;
  const bool fwd = dir == 0;
  if (transform == 1) { lm = (fwd || ordered) ? L_C_ORD : L_C_Z; sm = (fwd && !ordered) ? S_C_Z : S_C_ORD; }
  else if (fwd) { lm = L_R_TIME; sm = ordered ? S_R_ORD : S_R_Z; }
  else { lm = ordered ? L_R_ORD : L_R_Z; sm = S_R_TIME; }
  if (wmode && !tsw_plan_ok(Pn, A)) return -2;
  ts_build_stages<T>(P, Nc, Pn, A, tw_off, lm, sm, wmode);
  P.total_items = (unsigned)(batch * P.group_items);
  if (!wmode) P.total_items = (unsigned)((batch + (long long)(P.nstages - 1) * lag) * P.group_items);   // interleaved order
  if (wmode) {
    const std::vector<T> last = tsw_last_table<T>(ts_radix(A[Pn - 1]));
    const cpx<T>* twL = reinterpret_cast<const cpx<T>*>(last.data());
    return fwd ? tsw_emulate<T, -1>(P, twL, window, seed) : tsw_emulate<T, +1>(P, twL, window, seed);
  }
  return fwd ? ts_emulate<T, -1>(P, window, seed) : ts_emulate<T, +1>(P, window, seed);
}
// radices come from PFFFT_B200_TS_RADICES when set (ts_factorize reads it), else the default factorisation
extern "C" int emu_ts(int prec, int N, int transform, int dir, int ordered, const void* in, void* out, long long batch,
                      int lag, int window, unsigned seed) {
  if (prec == 0) return emu_ts_t<float>(N, transform, dir, ordered, (const float*)in, (float*)out, batch, lag, window, seed);
  return emu_ts_t<double>(N, transform, dir, ordered, (const double*)in, (double*)out, batch, lag, window, seed);
}
extern "C" int emu_tsw(int N, int transform, int dir, int ordered, const void* in, void* out, long long batch,
                       int lag, int window, unsigned seed) {
  return emu_ts_t<float>(N, transform, dir, ordered, (const float*)in, (float*)out, batch, lag, window, seed, true);
}
extern "C" int emu_ts_factorize(int Nc, int* P, int* A) { return ts_factorize(Nc, P, A) ? 1 : 0; }

// ---- compile-time-radix CTA kernels (radix_kernels.cuh): the stages of one transform stepped thread by thread
#include "../../pffft_b200/csrc/radix_kernels.cuh"
template <typename T, int R1, int R2, int R3, int LM, int SM, int SIGN>
static void radix_emulate(const T* in, T* out, int N, const cpx<T>* tw, const cpx<T>* twr) {
  using S = RadixShape<R1, R2, R3>;
  constexpr bool partner = (SM == S_R_ORD || SM == S_R_Z);
  std::vector<cpx<T>> buf(S::NCP);
  std::vector<T> o(2 * (size_t)S::NC);
  constexpr bool prerot = (LM == L_R_ORD || LM == L_R_Z);
  constexpr bool kPairsIn = radix_pairs_in_wanted<T, R1, R2, R3, LM, SIGN>();
  if (kPairsIn) {                                                // backward real: pre-rotation in registers, as in the kernel
    for (auto& e : buf) { e.x = (T)NAN; e.y = (T)NAN; }
    for (int li = 0; li < S::TT; ++li) radix_first_pairs<T, R1, R2, R3, LM>(li, in, N, twr, tw, buf.data());
  } else {
  if (prerot) for (int li = 0; li < S::TT; ++li) radix_prerotate<T, LM>(li, S::TT, in, N, S::NC, twr, buf.data());
  std::vector<std::vector<cpx<T>>> r1(S::TT, std::vector<cpx<T>>(R1));
  for (int li = 0; li < S::TT; ++li) {
    cpx<T> a[R1];
    if (prerot) radix_stage1_load<T, R1, R2, R3, LM, true>(li, in, N, twr, buf.data(), a);
    else radix_stage1_load<T, R1, R2, R3, LM, false>(li, in, N, twr, buf.data(), a);
    for (int j = 0; j < R1; ++j) r1[li][j] = a[j];
  }
  for (int li = 0; li < S::TT; ++li) {
    cpx<T> a[R1];
    for (int j = 0; j < R1; ++j) a[j] = r1[li][j];
    radix_stage1_store<T, R1, R2, R3, SIGN>(li, a, tw, buf.data());
    if (S::STAGES == 1) radix_emit<T, R1, S::M1, SM>(li, a, o.data(), N, buf.data());
  }
  }
  // forward real with a last radix <= 16: pair rotation in registers (radix_last_pairs), as in the kernel
  constexpr bool kPairs = radix_pairs_wanted<T, R1, R2, R3, SM, SIGN>();
  if (kPairs) for (auto& e : o) e = (T)NAN;
  if (S::STAGES == 2 && kPairs) {
    for (int li = 0; li < S::TT; ++li) radix_last_pairs<T, R2, S::M2, S::NC, SM, true, R1>(li, buf.data(), o.data(), N, twr);
  } else if (S::STAGES == 2) {
    std::vector<std::vector<cpx<T>>> regs(S::TT, std::vector<cpx<T>>(R2));
    for (int li = 0; li < S::TT; ++li) { cpx<T> a[R2]; radix_stage2_read<T, R1, R2, R3>(li, buf.data(), a); for (int j = 0; j < R2; ++j) regs[li][j] = a[j]; }
    for (int li = 0; li < S::M2; ++li) { cpx<T> a[R2]; for (int j = 0; j < R2; ++j) a[j] = regs[li][j]; dft_small<R2, SIGN>(a); radix_emit<T, R2, S::M2, SM>(li, a, o.data(), N, buf.data()); }
  } else if (S::STAGES == 3) {
    std::vector<std::vector<cpx<T>>> regs(S::TT, std::vector<cpx<T>>(R2));
    for (int li = 0; li < S::TT; ++li) { cpx<T> a[R2]; radix_stage2_read<T, R1, R2, R3>(li, buf.data(), a); for (int j = 0; j < R2; ++j) regs[li][j] = a[j]; }
    for (int li = 0; li < S::TT; ++li) { cpx<T> a[R2]; for (int j = 0; j < R2; ++j) a[j] = regs[li][j]; radix_stage2_write<T, R1, R2, R3, SIGN>(li, a, tw, buf.data()); }
    if (kPairs) {
      for (int li = 0; li < S::TT; ++li) radix_last_pairs<T, R3, S::M3, S::NC, SM, false, R1>(li, buf.data(), o.data(), N, twr);
    } else {
    std::vector<std::vector<cpx<T>>> r3(S::TT, std::vector<cpx<T>>(R3));
    for (int li = 0; li < S::TT; ++li) { cpx<T> c[R3]; radix_stage3_read<T, R1, R2, R3>(li, buf.data(), c); for (int j = 0; j < R3; ++j) r3[li][j] = c[j]; }
    for (int li = 0; li < S::M3; ++li) { cpx<T> c[R3]; for (int j = 0; j < R3; ++j) c[j] = r3[li][j]; dft_small<R3, SIGN>(c); radix_emit<T, R3, S::M3, SM>(li, c, o.data(), N, buf.data()); }
    }
  }
  if (partner && !kPairs) for (int k = 0; k < S::NC / 2; ++k) real_post_pair<SM, T>(o.data(), buf.data(), k, N, S::NC, twr);
  memcpy(out, o.data(), sizeof(T) * 2 * (size_t)S::NC);
}
template <int R1, int R2, int R3>
static int radix_emu_modes(int N, int lm, int sm, int sign, const float* in, float* out, const cf* tw, const cf* twr) {
#define PF_RX(L, S_, SG) if (lm == L && sm == S_ && sign == SG) { radix_emulate<float, R1, R2, R3, L, S_, SG>(in, out, N, tw, twr); return 0; }
  PF_RX(L_C_ORD, S_C_ORD, -1) PF_RX(L_C_ORD, S_C_ORD, +1) PF_RX(L_C_ORD, S_C_Z, -1) PF_RX(L_C_Z, S_C_ORD, +1)
  PF_RX(L_R_TIME, S_R_ORD, -1) PF_RX(L_R_TIME, S_R_Z, -1) PF_RX(L_R_ORD, S_R_TIME, +1) PF_RX(L_R_Z, S_R_TIME, +1)
#undef PF_RX
  return -2;
}
// double-precision cores (radix_d.cu)
template <int R1, int R2, int R3>
static int radix_emu_modes_d(int N, int lm, int sm, int sign, const double* in, double* out, const cd* tw, const cd* twr) {
#define PF_RX(L, S_, SG) if (lm == L && sm == S_ && sign == SG) { radix_emulate<double, R1, R2, R3, L, S_, SG>(in, out, N, tw, twr); return 0; }
  PF_RX(L_C_ORD, S_C_ORD, -1) PF_RX(L_C_ORD, S_C_ORD, +1) PF_RX(L_C_ORD, S_C_Z, -1) PF_RX(L_C_Z, S_C_ORD, +1)
  PF_RX(L_R_TIME, S_R_ORD, -1) PF_RX(L_R_TIME, S_R_Z, -1) PF_RX(L_R_ORD, S_R_TIME, +1) PF_RX(L_R_Z, S_R_TIME, +1)
#undef PF_RX
  return -2;
}
extern "C" int emu_radix_d(int N, int transform, int dir, int ordered, const double* in, double* out) {
  const int Nc = transform == 0 ? N / 2 : N;
  std::vector<double> tw(2 * (size_t)Nc), twr(2 * (size_t)(N / 2));
  pfplan::fill_roots<double>(tw.data(), Nc, Nc);
  pfplan::fill_roots<double>(twr.data(), N / 2, N);
  const cd* t1 = reinterpret_cast<const cd*>(tw.data()); const cd* t2 = reinterpret_cast<const cd*>(twr.data());
  int lm, sm; const bool fwd = dir == 0;
  if (transform == 1) { lm = (fwd || ordered) ? L_C_ORD : L_C_Z; sm = (fwd && !ordered) ? S_C_Z : S_C_ORD; }
  else if (fwd) { lm = L_R_TIME; sm = ordered ? S_R_ORD : S_R_Z; }
  else { lm = ordered ? L_R_ORD : L_R_Z; sm = S_R_TIME; }
  const int sign = fwd ? -1 : +1;
  switch (Nc) {
#define D(nc, r1, r2, r3) case nc: return radix_emu_modes_d<r1, r2, r3>(N, lm, sm, sign, in, out, t1, t2);
    D(16, 4, 4, 1) D(32, 8, 4, 1) D(48, 8, 6, 1) D(64, 8, 8, 1) D(80, 10, 8, 1) D(96, 12, 8, 1) D(128, 8, 4, 4) D(144, 12, 12, 1)
    D(160, 8, 5, 4) D(192, 8, 6, 4) D(240, 8, 6, 5) D(256, 8, 8, 4) D(288, 8, 6, 6) D(320, 8, 8, 5) D(384, 8, 8, 6) D(400, 10, 10, 4)
    D(432, 9, 8, 6) D(480, 10, 8, 6) D(1296, 12, 12, 9) D(2000, 10, 20, 10)
    D(576, 9, 8, 8) D(640, 10, 8, 8) D(720, 10, 9, 8) D(768, 12, 8, 8) D(800, 10, 10, 8) D(864, 12, 9, 8) D(960, 12, 10, 8)
    D(1152, 12, 12, 8) D(1200, 12, 10, 10) D(1280, 16, 10, 8) D(1440, 12, 12, 10) D(1600, 16, 10, 10) D(1728, 12, 12, 12) D(1920, 16, 12, 10)
    D(2160, 12, 12, 15) D(2304, 16, 12, 12) D(2400, 16, 15, 10) D(2560, 16, 16, 10) D(2592, 9, 16, 18) D(2880, 16, 15, 12)
    D(3456, 16, 18, 12) D(3600, 16, 15, 15) D(3840, 16, 16, 15)
#undef D
  }
  return -1;
}
extern "C" int emu_radix(int N, int transform, int dir, int ordered, const float* in, float* out) {
  const int Nc = transform == 0 ? N / 2 : N;
  std::vector<float> tw(2 * (size_t)Nc), twr(2 * (size_t)(N / 2));
  pfplan::fill_roots<float>(tw.data(), Nc, Nc);
  pfplan::fill_roots<float>(twr.data(), N / 2, N);
  const cf* t1 = reinterpret_cast<const cf*>(tw.data()); const cf* t2 = reinterpret_cast<const cf*>(twr.data());
  int lm, sm; const bool fwd = dir == 0;
  if (transform == 1) { lm = (fwd || ordered) ? L_C_ORD : L_C_Z; sm = (fwd && !ordered) ? S_C_Z : S_C_ORD; }
  else if (fwd) { lm = L_R_TIME; sm = ordered ? S_R_ORD : S_R_Z; }
  else { lm = ordered ? L_R_ORD : L_R_Z; sm = S_R_TIME; }
  const int sign = fwd ? -1 : +1;
  switch (Nc) {
    case 16: return radix_emu_modes<4, 4, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 48: return radix_emu_modes<8, 6, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 80: return radix_emu_modes<10, 8, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 144: return radix_emu_modes<12, 12, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 240: return radix_emu_modes<16, 15, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 400: return radix_emu_modes<20, 20, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 432: return radix_emu_modes<24, 18, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 720: return radix_emu_modes<30, 24, 1>(N, lm, sm, sign, in, out, t1, t2);
    case 1152: return radix_emu_modes<12, 12, 8>(N, lm, sm, sign, in, out, t1, t2);
    case 1200: return radix_emu_modes<12, 10, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 1280: return radix_emu_modes<16, 10, 8>(N, lm, sm, sign, in, out, t1, t2);
    case 1440: return radix_emu_modes<12, 12, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 1600: return radix_emu_modes<16, 10, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 1728: return radix_emu_modes<12, 12, 12>(N, lm, sm, sign, in, out, t1, t2);
    case 1920: return radix_emu_modes<16, 12, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 2304: return radix_emu_modes<16, 12, 12>(N, lm, sm, sign, in, out, t1, t2);
    case 3200: return radix_emu_modes<20, 16, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 3456: return radix_emu_modes<16, 18, 12>(N, lm, sm, sign, in, out, t1, t2);
    case 3600: return radix_emu_modes<16, 15, 15>(N, lm, sm, sign, in, out, t1, t2);
    case 3840: return radix_emu_modes<16, 16, 15>(N, lm, sm, sign, in, out, t1, t2);
    case 2160: return radix_emu_modes<12, 12, 15>(N, lm, sm, sign, in, out, t1, t2);
    case 2400: return radix_emu_modes<16, 15, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 2880: return radix_emu_modes<16, 15, 12>(N, lm, sm, sign, in, out, t1, t2);
    case 4320: return radix_emu_modes<16, 18, 15>(N, lm, sm, sign, in, out, t1, t2);
    case 4608: return radix_emu_modes<16, 16, 18>(N, lm, sm, sign, in, out, t1, t2);
    case 4800: return radix_emu_modes<16, 20, 15>(N, lm, sm, sign, in, out, t1, t2);
    case 5184: return radix_emu_modes<16, 18, 18>(N, lm, sm, sign, in, out, t1, t2);
    case 5760: return radix_emu_modes<16, 18, 20>(N, lm, sm, sign, in, out, t1, t2);
    case 6400: return radix_emu_modes<16, 20, 20>(N, lm, sm, sign, in, out, t1, t2);
    case 6912: return radix_emu_modes<16, 18, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 7200: return radix_emu_modes<15, 20, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 8000: return radix_emu_modes<20, 20, 20>(N, lm, sm, sign, in, out, t1, t2);
    case 7680: return radix_emu_modes<16, 20, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 9216: return radix_emu_modes<16, 24, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 2560: return radix_emu_modes<16, 16, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 9600: return radix_emu_modes<20, 20, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 10800: return radix_emu_modes<18, 20, 30>(N, lm, sm, sign, in, out, t1, t2);
    case 11520: return radix_emu_modes<20, 24, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 12960: return radix_emu_modes<18, 24, 30>(N, lm, sm, sign, in, out, t1, t2);
    case 13824: return radix_emu_modes<24, 24, 24>(N, lm, sm, sign, in, out, t1, t2);
    case 14400: return radix_emu_modes<24, 24, 25>(N, lm, sm, sign, in, out, t1, t2);
    case 5120: return radix_emu_modes<16, 16, 20>(N, lm, sm, sign, in, out, t1, t2);
    case 1296: return radix_emu_modes<12, 12, 9>(N, lm, sm, sign, in, out, t1, t2);
    case 2000: return radix_emu_modes<25, 10, 8>(N, lm, sm, sign, in, out, t1, t2);
    case 2592: return radix_emu_modes<9, 16, 18>(N, lm, sm, sign, in, out, t1, t2);
    case 4000: return radix_emu_modes<25, 16, 10>(N, lm, sm, sign, in, out, t1, t2);
    case 6000: return radix_emu_modes<15, 20, 20>(N, lm, sm, sign, in, out, t1, t2);
    case 12000: return radix_emu_modes<25, 24, 20>(N, lm, sm, sign, in, out, t1, t2);
  }
  return -1;
}
