// ts_plan.h -- host-side plan algebra of the tiled Stockham pipeline (no CUDA calls): factorisation of a complex core into
// radices 16*A, work items per pass, the stage list of a call, and the per-radix twiddle tables.  Shared by ts.cu and the
// CPU stepping harness (tests/emu).
#pragma once
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "plan.h"
#include "tsw_kernels.cuh"

namespace pf {

// work items per pass
// (wmode: the warp-sized work items of tsw_kernels.cuh)
inline int ts_tiles(int Nc, int A, bool wmode = false) {
  if (A >= kTsSmallFlag) return (Nc / (A - kTsSmallFlag) + 255) / 256;
  const int m = Nc / (16 * A), cols = wmode ? tsw_cols_for(A) : ts_cols_for(A);
  return (m + cols - 1) / cols;
}
inline int ts_radix(int A) { return A >= kTsSmallFlag ? A - kTsSmallFlag : 16 * A; }
static const int kTsRadixA[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2, 1};

// rest = prod of P factors from kTsRadixA, as balanced as possible (smallest maximum, then largest minimum), non-increasing
inline bool ts_search(long long rest, int depth, int P, int amax, int* cur, int* best, int* best_max, int* best_min) {
  if (depth == P) {
    if (rest != 1) return false;
    int mx = 0, mn = 99;
    for (int i = 0; i < P; ++i) { if (cur[i] > mx) mx = cur[i]; if (cur[i] < mn) mn = cur[i]; }
    if (mx < *best_max || (mx == *best_max && mn > *best_min)) { *best_max = mx; *best_min = mn; memcpy(best, cur, sizeof(int) * P); }
    return true;
  }
  bool any = false;
  for (int a : kTsRadixA) {
    if (a > amax || rest % a) continue;
    cur[depth] = a;
    any |= ts_search(rest / a, depth + 1, P, a, cur, best, best_max, best_min);
  }
  return any;
}
// Nc = prod of `full` radices 16*A_i (full >= 2, larger first), optionally closed by one small radix A_s (2..15, no factor
// 16: for cores with too few factors of two).  The plan with the fewest passes wins.  A_out[i] >= kTsSmallFlag marks the
// small pass.  false: no such factorisation.
inline bool ts_factorize(int Nc, int* P_out, int* A_out) {
  if (const char* e = getenv("PFFFT_B200_TS_RADICES")) {          // "256,256,16" or "240,160,s10": explicit radices (tuning / tests)
    int P = 0, A[4]; long long prod = 1; const char* p = e;
    while (*p && P < 4) {
      const bool small = (*p == 's');
      if (small) ++p;
      const int r = atoi(p);
      bool ok = false;
      if (small) { for (int a : kTsRadixA) ok |= (r == a && a > 1 && a < 16); }
      else { for (int a : kTsRadixA) ok |= (r == 16 * a); }
      if (!ok) { P = 0; break; }
      A[P++] = small ? r + kTsSmallFlag : r / 16; prod *= r;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    bool shape = P >= 2 && prod == Nc;
    for (int i = 0; shape && i < P; ++i) if (A[i] >= kTsSmallFlag && (i != P - 1 || P < 3)) shape = false;
    if (shape) { *P_out = P; memcpy(A_out, A, sizeof(int) * P); return true; }
  }
  for (int total = 2; total <= 4; ++total) {
    for (int small = 0; small <= 1; ++small) {
      const int full = total - small;
      if (full < 2) continue;
      long long unit = 1;
      for (int i = 0; i < full; ++i) unit *= 16;
      if (Nc % unit) continue;
      const long long rest = Nc / unit;
      if (!small) {
        int cur[4], best[4], bmax = 99, bmin = 0;
        if (ts_search(rest, 0, full, 16, cur, best, &bmax, &bmin)) { *P_out = full; memcpy(A_out, best, sizeof(int) * full); return true; }
      } else {
        for (int as : {2, 3, 4, 5, 6, 8, 9, 10, 12, 15}) {           // smallest closing radix first
          if (rest % as) continue;
          int cur[4], best[4], bmax = 99, bmin = 0;
          if (ts_search(rest / as, 0, full, 16, cur, best, &bmax, &bmin)) {
            memcpy(A_out, best, sizeof(int) * full); A_out[full] = as + kTsSmallFlag; *P_out = total; return true;
          }
        }
      }
    }
  }
  return false;
}

// per-radix tables [k_a*A + q] = exp(-2 pi i q k_a / (16 A)), one after the other; tw_off[i] = offset of pass i
template <typename T> inline std::vector<T> ts_radix_tables(int P, const int* A, int* tw_off) {
  size_t total = 0;
  for (int i = 0; i < P; ++i) { tw_off[i] = (int)total; if (A[i] < kTsSmallFlag) total += 16 * (size_t)A[i]; }
  std::vector<T> host(2 * total);
  for (int i = 0; i < P; ++i) {
    if (A[i] >= kTsSmallFlag) continue;
    const int a = A[i], R = 16 * a;
    for (int ka = 0; ka < 16; ++ka)
      for (int q = 0; q < a; ++q) {
        long double c, s;
        pfplan::unit_root((long long)q * ka, R, &c, &s);
        host[2 * (tw_off[i] + ka * a + q)] = (T)c; host[2 * (tw_off[i] + ka * a + q) + 1] = (T)s;
      }
  }
  return host;
}

// warp-sized work items (tsw_kernels.cuh): every pass a power-of-two radix 32 ... 256, no closing small pass
inline bool tsw_plan_ok(int P, const int* A) {
  for (int i = 0; i < P; ++i) if (!tsw_radix_ok(A[i])) return false;
  return P >= 2;
}
// exponent-indexed table of the last pass: exp(-2 pi i e / R), e < R
template <typename T> inline std::vector<T> tsw_last_table(int R) {
  std::vector<T> host(2 * (size_t)R);
  for (int e = 0; e < R; ++e) {
    long double c, s;
    pfplan::unit_root(e, R, &c, &s);
    host[2 * e] = (T)c; host[2 * e + 1] = (T)s;
  }
  return host;
}

// stage list of one call: [pre-rotation / z gather] + P passes + [post-rotation / z scatter]; sets nstages, group_items
template <typename T> inline void ts_build_stages(TsParams<T>& Q, int Nc, int P, const int* A, const int* tw_off, int lm, int sm,
                                                  bool wmode = false) {
  const bool pre = lm == L_C_Z || lm == L_R_ORD || lm == L_R_Z;
  const bool post = sm == S_C_Z || sm == S_R_ORD || sm == S_R_Z;
  const int chunks = (Nc + kTsChunk - 1) / kTsChunk;
  int ns = 0, ring_next = 0, cur_src = 0;
  if (pre) { Q.st[ns++] = TsStage{TS_PRE, 0, lm, chunks, 0, 2 + ring_next, 0, 0, 0}; cur_src = 2 + ring_next++; }
  int prod = 1;
  for (int i = 0; i < P; ++i) {
    const int R = ts_radix(A[i]);
    const bool last = (i == P - 1), small = A[i] >= kTsSmallFlag;
    const int dst = (last && !post) ? 1 : 2 + ring_next++;
    Q.st[ns++] = TsStage{small ? TS_SMALL : (i == 0 ? TS_FIRST : TS_LATER), small ? A[i] - kTsSmallFlag : A[i], 0, ts_tiles(Nc, A[i], wmode),
                         cur_src, dst, Nc / R, prod, tw_off[i]};
    cur_src = dst; prod *= R;
  }
  if (post) Q.st[ns++] = TsStage{TS_POST, 0, sm, chunks, cur_src, 1, 0, 0, 0};
  Q.nstages = ns;
  long long group = 0;
  for (int i = 0; i < ns; ++i) group += Q.st[i].tiles;
  Q.group_items = (int)group;
}

}  // namespace pf
