// emu_fast.cuh -- CPU stepping of the size-tuned kernels' PF_HD phases (development harness)
#pragma once
#include <vector>
#include "../../pffft_b200/csrc/fast_kernels.cuh"
#include "../../pffft_b200/csrc/plan.h"

template <int SIGN> static void emu_w1024_run(const float* in, float* out, long long batch) {
  using namespace pf;
  std::vector<cf> tw(1024);
  for (int k2 = 0; k2 < 32; ++k2) for (int n1 = 0; n1 < 32; ++n1) {
    long double c, s; pfplan::unit_root((long long)n1 * k2, 1024, &c, &s);
    tw[k2 * 32 + n1] = mk<float>((float)c, (float)s);
  }
  std::vector<cf> tile(kW1024Tile);
  for (long long t = 0; t < batch; ++t) {
    const cf* src = reinterpret_cast<const cf*>(in) + t * 1024;
    cf* dst = reinterpret_cast<cf*>(out) + t * 1024;
    for (int lane = 0; lane < 32; ++lane) {
      cf v[32];
      for (int p = 0; p < 32; ++p) v[p] = src[lane + 32 * brev5(p)];
      w1024_rows<SIGN>(v, lane, tw.data(), tile.data());
    }
    for (int lane = 0; lane < 32; ++lane) {
      cf v[32];
      w1024_cols<SIGN>(v, lane, tile.data());
      for (int k1 = 0; k1 < 32; ++k1) dst[lane + 32 * k1] = v[k1];
    }
  }
}
extern "C" int emu_w1024(int dir, const float* in, float* out, long long batch) {
  if (dir == 0) emu_w1024_run<-1>(in, out, batch); else emu_w1024_run<+1>(in, out, batch);
  return 0;
}

// ---- CTA kernel (cta_kernels.cuh): phases stepped thread by thread
#include "../../pffft_b200/csrc/cta_kernels.cuh"
template <int C, int LM, int SM, int SIGN>
static void emu_k2_run(const float* in, float* out, int N, long long in_limit, int out_count) {
  using namespace pf;
  using K = K2<C>;
  const int Nc = K::NC;
  std::vector<float> tw1(2 * (size_t)Nc), tw2(2 * 16 * C), twr(2 * (size_t)(N / 2 > 0 ? N / 2 : 1));
  for (int ka = 0; ka < 16; ++ka) for (int m = 0; m < K::BC; ++m) {
    long double c, s; pfplan::unit_root((long long)m * ka, Nc, &c, &s);
    tw1[2 * (ka * K::BC + m)] = (float)c; tw1[2 * (ka * K::BC + m) + 1] = (float)s;
  }
  for (int kb = 0; kb < 16; ++kb) for (int nc = 0; nc < C; ++nc) {
    long double c, s; pfplan::unit_root((long long)nc * kb, K::BC, &c, &s);
    tw2[2 * (kb * C + nc)] = (float)c; tw2[2 * (kb * C + nc) + 1] = (float)s;
  }
  pfplan::fill_roots<float>(twr.data(), N / 2, N);
  const cf* ptw1 = reinterpret_cast<const cf*>(tw1.data());
  const cf* ptw2 = reinterpret_cast<const cf*>(tw2.data());
  const cf* ptwr = reinterpret_cast<const cf*>(twr.data());
  std::vector<cf> tile(Nc), nat(Nc);
  if constexpr (LM == L_R_ORD && SM == S_R_TIME && SIGN > 0 && C <= 8) {      // mirrored (decimation in frequency) backward real
    for (auto& e : tile) { e.x = NAN; e.y = NAN; }
    for (int t = 0; t < K::T; ++t) k2b_pass1_pairs<C, float>(t, in, N, ptwr, ptw2, tile.data());
    for (int t = 0; t < K::T; ++t) k2b_pass2<C, float>(t, ptw1, tile.data());
    for (int t = 0; t < K::T; ++t) {
      cf v[16];
      k2b_pass3<C, float>(t, tile.data(), v);
      for (int na = 0; na < 16; ++na) store_elem<SM, float>(out, t + K::BC * na, v[na], N, out_count, vec_aligned<float>(out));
    }
    return;
  }
  for (int t = 0; t < K::T; ++t) k2_pass1<C, LM, SIGN, false, float>(t, in, N, ptwr, in_limit, vec_aligned<float>(in), ptw1, tile.data());
  for (int t = 0; t < K::T; ++t) k2_pass2<C, SIGN, float>(t, ptw2, tile.data());
  constexpr bool partner = (SM == S_R_ORD || SM == S_R_Z);
  if (partner && C <= 8) {
    for (int t = 0; t < K::T; ++t) { cf u[16]; k2_pass3_pairs<C, SIGN, float>(t, tile.data(), u); k2_store_pairs<C, SM, float>(t, u, out, N, ptwr); }
    return;
  }
  for (int t = 0; t < K::T; ++t) {
    cf u[16];
    k2_pass3<C, SIGN, float>(t, tile.data(), u);
    for (int r = 0; r < 16 / C; ++r) for (int kc = 0; kc < C; ++kc) {
      const int k = k2_out_index<C>(t, r, kc);
      if (partner) nat[k] = u[r * C + kc];
      else store_elem<SM, float>(out, k, u[r * C + kc], N, out_count, vec_aligned<float>(out));
    }
  }
  if (partner)
    for (int t = 0; t < K::T; ++t) for (int j = 0; j < 8; ++j)
      real_post_pair<SM, float>(out, nat.data(), t + K::T * j, N, Nc, ptwr);
}
template <int C>
static int emu_k2_c(int N, int lm, int sm, int dir, const float* in, float* out, long long in_limit, int out_count) {
  using namespace pf;
#define K2CASE(L, S, SG) if (lm == L && sm == S && dir == (SG > 0 ? 1 : 0)) { emu_k2_run<C, L, S, SG>(in, out, N, in_limit, out_count); return 0; }
  K2CASE(L_C_ORD, S_C_ORD, -1) K2CASE(L_C_ORD, S_C_Z, -1) K2CASE(L_C_ORD, S_C_ORD, +1) K2CASE(L_C_Z, S_C_ORD, +1)
  K2CASE(L_R_TIME, S_R_ORD, -1) K2CASE(L_R_TIME, S_R_Z, -1) K2CASE(L_R_ORD, S_R_TIME, +1) K2CASE(L_R_Z, S_R_TIME, +1)
#undef K2CASE
  return -2;
}
// Nc = complex core length (512..4096), N = API length (Nc complex, 2*Nc real)
extern "C" int emu_k2(int Nc, int N, int lm, int sm, int dir, const float* in, float* out, long long in_limit, int out_count) {
  switch (Nc) {
    case 512: return emu_k2_c<2>(N, lm, sm, dir, in, out, in_limit, out_count);
    case 1024: return emu_k2_c<4>(N, lm, sm, dir, in, out, in_limit, out_count);
    case 2048: return emu_k2_c<8>(N, lm, sm, dir, in, out, in_limit, out_count);
    case 4096: return emu_k2_c<16>(N, lm, sm, dir, in, out, in_limit, out_count);
  }
  return -1;
}
// bank-conflict audit of the swizzled tile: worst number of distinct-address hits per bank-pair over all
// half-warps of the three passes (1 = conflict free)
template <int C> static int emu_k2_conflicts_c() {
  using K = pf::K2<C>;
  int worst = 1;
  auto audit = [&](auto addr_of_thread, int nthreads) {
    for (int h = 0; h < nthreads; h += 16) {
      int cnt[16] = {0};
      for (int l = 0; l < 16 && h + l < nthreads; ++l) cnt[addr_of_thread(h + l) & 15]++;
      for (int b = 0; b < 16; ++b) if (cnt[b] > worst) worst = cnt[b];
    }
  };
  for (int ka = 0; ka < 16; ++ka) audit([&](int m) { return K::idx(ka, m / C, m % C); }, K::T);                 // pass 1 writes
  for (int jb = 0; jb < 16; ++jb) audit([&](int t) { return K::idx(t / C, jb, t % C); }, K::T);                  // pass 2 r/w
  for (int r = 0; r < 16 / C; ++r) for (int jc = 0; jc < C; ++jc)
    audit([&](int t) { return K::idx(t & 15, (t >> 4) + C * r, jc); }, K::T);                                    // pass 3 reads
  return worst;
}
extern "C" int emu_k2_conflicts(int C) {
  switch (C) { case 2: return emu_k2_conflicts_c<2>(); case 4: return emu_k2_conflicts_c<4>();
               case 8: return emu_k2_conflicts_c<8>(); case 16: return emu_k2_conflicts_c<16>(); }
  return -1;
}

// ---- fused overlap-save kernel (fastconv_kernels.cuh), stepped block by block
#include "../../pffft_b200/csrc/fastconv_kernels.cuh"
template <int C>
static long long emu_fastconv_c(const float* x, long long len, const float* h, int taps, int flush, float* y) {
  using namespace pf;
  using K = K2<C>;
  const int Nc = K::NC, Nfft = 2 * Nc;
  std::vector<float> tw1(2 * (size_t)Nc), tw2(2 * 16 * C), twr(2 * (size_t)Nc);
  for (int ka = 0; ka < 16; ++ka) for (int m = 0; m < K::BC; ++m) {
    long double c, s; pfplan::unit_root((long long)m * ka, Nc, &c, &s);
    tw1[2 * (ka * K::BC + m)] = (float)c; tw1[2 * (ka * K::BC + m) + 1] = (float)s;
  }
  for (int kb = 0; kb < 16; ++kb) for (int nc = 0; nc < C; ++nc) {
    long double c, s; pfplan::unit_root((long long)nc * kb, K::BC, &c, &s);
    tw2[2 * (kb * C + nc)] = (float)c; tw2[2 * (kb * C + nc) + 1] = (float)s;
  }
  pfplan::fill_roots<float>(twr.data(), Nc, Nfft);
  const cf* ptw1 = reinterpret_cast<const cf*>(tw1.data());
  const cf* ptw2 = reinterpret_cast<const cf*>(tw2.data());
  const cf* ptwr = reinterpret_cast<const cf*>(twr.data());
  // filter spectrum, canonical layout (time-reversed taps placed circularly, ref pffastconv.c:99-106)
  std::vector<float> ht(Nfft, 0.f), Hc(Nfft, 0.f);
  for (int i = 0; i < taps; ++i) ht[(Nfft - i) & (Nfft - 1)] = h[taps - 1 - i];
  emu_k2_run<C, L_R_TIME, S_R_ORD, -1>(ht.data(), Hc.data(), Nfft, -1, Nfft);
  const pfplan::BlockPlan bp = pfplan::plan_blocks(len, Nfft, taps, flush, false);
  const long long nblk = bp.n_full + (bp.tail_off >= 0 ? 1 : 0);
  std::vector<cf> tile(Nc), nat(Nc);
  const float scale = (float)(1.0 / Nfft);
  for (long long b = 0; b < nblk; ++b) {
    const long long off = b * bp.stride;
    const int out_count = b < bp.n_full ? bp.stride : bp.tail_out;
    const float* ibase = x + off; float* obase = y + off;
    for (int t = 0; t < K::T; ++t) k2_pass1<C, L_R_TIME, -1, false, float>(t, ibase, Nfft, ptwr, len - off, vec_aligned<float>(ibase), ptw1, tile.data());
    for (int t = 0; t < K::T; ++t) k2_pass2<C, -1, float>(t, ptw2, tile.data());
    for (int t = 0; t < K::T; ++t) { cf u[16]; k2_pass3<C, -1, float>(t, tile.data(), u);
      for (int r = 0; r < 16 / C; ++r) for (int kc = 0; kc < C; ++kc) nat[k2_out_index<C>(t, r, kc)] = u[r * C + kc]; }
    for (int t = 0; t < K::T; ++t) for (int j = 0; j < 8; ++j)
      fastconv_pair<float>(nat.data(), t + K::T * j, Nc, ptwr, reinterpret_cast<const cf*>(Hc.data()), scale);
    for (int t = 0; t < K::T; ++t) k2_pass1_smem<C, +1, float>(t, nat.data(), ptw1, tile.data());
    for (int t = 0; t < K::T; ++t) k2_pass2<C, +1, float>(t, ptw2, tile.data());
    for (int t = 0; t < K::T; ++t) { cf u[16]; k2_pass3<C, +1, float>(t, tile.data(), u);
      for (int r = 0; r < 16 / C; ++r) for (int kc = 0; kc < C; ++kc)
        store_elem<S_R_TIME, float>(obase, k2_out_index<C>(t, r, kc), u[r * C + kc], Nfft, out_count, vec_aligned<float>(obase)); }
  }
  return bp.produced;
}
extern "C" long long emu_fastconv(int Nfft, const float* x, long long len, const float* h, int taps, int flush, float* y) {
  switch (Nfft) {
    case 1024: return emu_fastconv_c<2>(x, len, h, taps, flush, y);
    case 2048: return emu_fastconv_c<4>(x, len, h, taps, flush, y);
    case 4096: return emu_fastconv_c<8>(x, len, h, taps, flush, y);
    case 8192: return emu_fastconv_c<16>(x, len, h, taps, flush, y);
  }
  return -1;
}

// ---- small warp kernel phases (N = 32..256 complex)
template <int R2, int SIGN> static void emu_wsmall_run(const float* in, float* out, long long batch) {
  using namespace pf;
  constexpr int NC = 32 * R2, TW = 32 / R2;
  std::vector<cf> tw(NC > 32 ? NC : 32);
  for (int k2 = 0; k2 < R2; ++k2) for (int l = 0; l < 32; ++l) {
    long double c, s; pfplan::unit_root((long long)l * k2, NC, &c, &s);
    tw[k2 * 32 + l] = mk<float>((float)c, (float)s);
  }
  std::vector<cf> tile(kW1024Tile);
  const long long nchunks = (batch + TW - 1) / TW;
  for (long long c = 0; c < nchunks; ++c) {
    const cf* src = reinterpret_cast<const cf*>(in) + c * 1024;
    cf* dst = reinterpret_cast<cf*>(out) + c * 1024;
    const long long left = batch - c * TW; const int nvalid = left >= TW ? TW : (int)left;
    for (int lane = 0; lane < 32; ++lane) {
      cf v[32];
      for (int j = 0; j < TW; ++j) for (int p = 0; p < R2; ++p)
        v[j * R2 + p] = (j < nvalid) ? src[lane + 32 * (j * R2 + brevR2<R2>(p))] : mk<float>(0.f, 0.f);
      wsmall_rows<R2, SIGN>(v, lane, tw.data(), tile.data());
    }
    for (int lane = 0; lane < 32; ++lane) {
      cf v[32];
      w1024_cols<SIGN>(v, lane, tile.data());
      const int j = lane / R2, k2 = lane % R2;
      if (j < nvalid) for (int k1 = 0; k1 < 32; ++k1) dst[j * NC + k2 + R2 * k1] = v[k1];
    }
  }
}
extern "C" int emu_wsmall(int N, int dir, const float* in, float* out, long long batch) {
#define WS(n, r) if (N == n) { if (dir == 0) emu_wsmall_run<r, -1>(in, out, batch); else emu_wsmall_run<r, +1>(in, out, batch); return 0; }
  WS(32, 1) WS(64, 2) WS(128, 4) WS(256, 8)
#undef WS
  return -1;
}

// ---- non-power-of-two warp kernel phases (N = 32*R2)
template <int R2, int SIGN> static void emu_wmixed_run(const float* in, float* out, long long batch) {
  using namespace pf;
  constexpr int NC = 32 * R2, TW = 32 / R2, COLS = TW * R2;
  std::vector<cf> tw(NC);
  for (int k2 = 0; k2 < R2; ++k2) for (int l = 0; l < 32; ++l) {
    long double c, s; pfplan::unit_root((long long)l * k2, NC, &c, &s);
    tw[k2 * 32 + l] = mk<float>((float)c, (float)s);
  }
  std::vector<cf> tile(kW1024Tile);
  const long long nchunks = (batch + TW - 1) / TW;
  for (long long c = 0; c < nchunks; ++c) {
    const cf* src = reinterpret_cast<const cf*>(in) + c * (TW * NC);
    cf* dst = reinterpret_cast<cf*>(out) + c * (TW * NC);
    const long long left = batch - c * TW; const int nvalid = left >= TW ? TW : (int)left;
    for (int lane = 0; lane < 32; ++lane) {
      cf v[32];
      for (int m = 0; m < 32; ++m) v[m] = mk<float>(0.f, 0.f);
      for (int j = 0; j < TW; ++j) for (int n2 = 0; n2 < R2; ++n2) if (j < nvalid) v[j * R2 + n2] = src[j * NC + lane + 32 * n2];
      wmixed_rows<R2, SIGN>(v, lane, tw.data(), tile.data());
    }
    std::vector<cf> res((size_t)TW * NC);
    for (int lane = 0; lane < COLS; ++lane) {
      cf v[32];
      w1024_cols<SIGN>(v, lane, tile.data());
      const int j = lane / R2, k2 = lane % R2;
      if (j < nvalid) for (int k1 = 0; k1 < 32; ++k1) dst[j * NC + k2 + R2 * k1] = v[k1];
    }
  }
}
extern "C" int emu_wmixed(int N, int dir, const float* in, float* out, long long batch) {
#define WM(r) if (N == 32 * r) { if (dir == 0) emu_wmixed_run<r, -1>(in, out, batch); else emu_wmixed_run<r, +1>(in, out, batch); return 0; }
  WM(3) WM(5) WM(6) WM(9) WM(10) WM(12) WM(15) WM(18) WM(20) WM(24) WM(25) WM(27) WM(30)
#undef WM
  return -1;
}

// ---- cluster kernel (cluster_kernels.cuh): the CL CTAs of one cluster stepped phase by phase, barriers = phase ends
#include "../../pffft_b200/csrc/cluster_kernels.cuh"
#include "../../pffft_b200/csrc/cta_hooks.cuh"   // split_fill_tables
struct EmuRemote { pf::cf* const* bases; PF_HD pf::cf* operator()(int owner) const { return bases[owner]; } };
template <int C, int CL, int Q, bool SCATTER, int SIGN>
static void emu_cluster_run(const float* in, float* out) {
  using namespace pf;
  using K = K2<C>; using G = KCL<C, CL, Q>;
  std::vector<float> tab(2 * split_table_cpx(G::NC, G::N2));
  split_fill_tables<float>(G::NC, G::N2, tab.data());
  const cf* tw1 = reinterpret_cast<const cf*>(tab.data());
  const cf* tw2 = tw1 + G::N2;
  const cf* twP = tw1 + cta_table_cpx(G::N2);
  std::vector<std::vector<cf>> tile(CL, std::vector<cf>(G::N2)), park(CL, std::vector<cf>((size_t)Q * G::N2));
  cf* bases[CL];
  for (int c = 0; c < CL; ++c) bases[c] = park[c].data();
  const EmuRemote remote{bases};
  const cf* src = reinterpret_cast<const cf*>(in);
  cf* dst = reinterpret_cast<cf*>(out);
  std::vector<cf> U((size_t)CL * K::T * 16);
  auto u_of = [&](int c, int t) -> cf (&)[16] { return *reinterpret_cast<cf(*)[16]>(&U[((size_t)c * K::T + t) * 16]); };
  if (SCATTER) {
    for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) cl_scatter<C, CL, float>(t, c, src, remote);
    for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) k2_pass1_smem<C, SIGN, float>(t, park[c].data(), tw1, tile[c].data());
    for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) k2_pass2<C, SIGN, float>(t, tw2, tile[c].data());
    for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) k2_pass3<C, SIGN, float>(t, tile[c].data(), u_of(c, t));
    for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) cl_park<C, CL, Q, SIGN, float>(t, c, u_of(c, t), twP, remote);
  } else {
    for (int q = 0; q < Q; ++q)
      for (int c = 0; c < CL; ++c) {
        const int n1 = c + CL * q;
        for (int t = 0; t < K::T; ++t) cl_row_issue<C, G::R, float>(t, src + n1, tile[c].data());
        for (int t = 0; t < K::T; ++t) cl_row_pass1<C, SIGN, float>(t, tw1, tile[c].data());
        for (int t = 0; t < K::T; ++t) k2_pass2<C, SIGN, float>(t, tw2, tile[c].data());
        for (int t = 0; t < K::T; ++t) { k2_pass3<C, SIGN, float>(t, tile[c].data(), u_of(c, t)); cl_park<C, CL, Q, SIGN, float>(t, n1, u_of(c, t), twP, remote); }
      }
  }
  for (int c = 0; c < CL; ++c) for (int t = 0; t < K::T; ++t) cl_combine<C, CL, Q, SIGN, float>(t, c, park[c].data(), dst);
}
extern "C" int emu_cluster(int CL, int Q, int scatter, int dir, const float* in, float* out) {
#define CLX(cl, q, sc) if (CL == cl && Q == q && (scatter != 0) == sc) { if (dir == 0) emu_cluster_run<16, cl, q, sc, -1>(in, out); else emu_cluster_run<16, cl, q, sc, +1>(in, out); return 0; }
  CLX(2, 1, false) CLX(2, 1, true) CLX(4, 1, false) CLX(4, 1, true) CLX(8, 1, false) CLX(8, 1, true) CLX(8, 2, false) CLX(16, 1, false) CLX(16, 1, true) CLX(4, 2, false) CLX(4, 4, false)
#undef CLX
  return -1;
}
// single-CTA two-level kernel (k_cta_split) with the row-major twiddle table
template <int C, int R, int SIGN>
static void emu_cta_split_run(const float* in, float* out) {
  using namespace pf;
  using K = K2<C>;
  constexpr int N2 = K::NC, Nc = R * N2;
  std::vector<float> tab(2 * split_table_cpx(Nc, N2));
  split_fill_tables<float>(Nc, N2, tab.data());
  const cf* tw1 = reinterpret_cast<const cf*>(tab.data());
  const cf* tw2 = tw1 + N2;
  const cf* twP = tw1 + cta_table_cpx(N2);
  std::vector<cf> tile(N2), rows((size_t)R * N2);
  const cf* src = reinterpret_cast<const cf*>(in);
  cf* dst = reinterpret_cast<cf*>(out);
  for (int n1 = 0; n1 < R; ++n1) {
    for (int t = 0; t < K::T; ++t) k2_pass1<C, L_C_ORD, SIGN, false, float>(t, reinterpret_cast<const float*>(src + n1), N2, nullptr, -1, true, tw1, tile.data(), R);
    for (int t = 0; t < K::T; ++t) k2_pass2<C, SIGN, float>(t, tw2, tile.data());
    std::vector<cf> U((size_t)K::T * 16);
    for (int t = 0; t < K::T; ++t) k2_pass3<C, SIGN, float>(t, tile.data(), *reinterpret_cast<cf(*)[16]>(&U[(size_t)t * 16]));
    for (int t = 0; t < K::T; ++t) split_park<C, R, SIGN, N2, 1, float>(t, n1, *reinterpret_cast<cf(*)[16]>(&U[(size_t)t * 16]), twP, rows.data());
  }
  for (int t = 0; t < K::T; ++t) split_combine_cols<C, R, SIGN, N2, 1, float>(t, rows.data(), dst);
}
extern "C" int emu_cta_split(int C, int R, int dir, const float* in, float* out) {
#define CSX(c, r) if (C == c && R == r) { if (dir == 0) emu_cta_split_run<c, r, -1>(in, out); else emu_cta_split_run<c, r, +1>(in, out); return 0; }
  CSX(16, 2) CSX(8, 3) CSX(4, 9) CSX(2, 15) CSX(8, 6)
#undef CSX
  return -1;
}

// ---- tiled two-dimensional large-N plan (tiled2d_kernels.cuh): pass A tiles then pass C tiles, thread by thread
#include "../../pffft_b200/csrc/tiled2d_kernels.cuh"
template <int A1, int A2, int SIGN>
static void emu_t2d_run(const float* in, float* out) {
  using namespace pf;
  using G = T2D<A1, A2>;
  std::vector<float> tab(2 * ((size_t)G::N1 + G::N2 + G::NC));
  t2d_fill_tables<float, A1, A2>(tab.data());
  const cf* twA = reinterpret_cast<const cf*>(tab.data());
  const cf* twC = twA + G::N2;
  const cf* tw2d = twC + G::N1;
  std::vector<cf> S(G::NC), tileA(16 * G::N2), tileC(16 * G::N1);
  const cf* x = reinterpret_cast<const cf*>(in);
  cf* X = reinterpret_cast<cf*>(out);
  for (int c = 0; c < G::N1 / 16; ++c) {
    for (int t = 0; t < G::TA; ++t) t2d_A1<A1, A2, SIGN, float>(t, x + 16 * c, twA, tileA.data());
    for (int t = 0; t < G::TA; ++t) t2d_A2<A1, A2, SIGN, float>(t, c, tileA.data(), tw2d, T2DScratchSink<float, G::N1>{S.data()});
  }
  for (int d = 0; d < G::N2 / 16; ++d) {
    for (int t = 0; t < G::TC; ++t) t2d_C1<A1, A2, SIGN, float>(t, S.data() + (size_t)16 * d * G::N1, twC, tileC.data());
    for (int t = 0; t < G::TC; ++t) t2d_C2<A1, A2, SIGN, float>(t, tileC.data(), X + 16 * d);
  }
}
extern "C" int emu_t2d(int A1, int A2, int dir, const float* in, float* out) {
#define T2X(a1, a2) if (A1 == a1 && A2 == a2) { if (dir == 0) emu_t2d_run<a1, a2, -1>(in, out); else emu_t2d_run<a1, a2, +1>(in, out); return 0; }
  T2X(8, 8) T2X(16, 8) T2X(8, 16) T2X(16, 16)
#undef T2X
  return -1;
}
// worst number of lanes of a half-warp that hit the same 8-byte bank pair in the exchange tiles (1 = conflict free)
template <int A1, int A2> static int emu_t2d_conflicts_c() {
  using namespace pf;
  int worst = 1;
  auto audit = [&](auto addr, int nthreads) {
    for (int h = 0; h < nthreads; h += 16) { int cnt[16] = {0}; for (int l = 0; l < 16; ++l) cnt[addr(h + l) & 15]++; for (int b = 0; b < 16; ++b) if (cnt[b] > worst) worst = cnt[b]; }
  };
  for (int ka = 0; ka < 16; ++ka) audit([&](int t) { return (ka * A2 + (t >> 4)) * 16 + (t & 15); }, 16 * A2);                       // A1 writes
  for (int q = 0; q < A2; ++q) audit([&](int t) { return (((t >> 4)) * A2 + q) * 16 + (t & 15); }, 16 * A2);                          // A2 reads
  for (int ka = 0; ka < 16; ++ka) audit([&](int t) { return (ka * A1 + t % A1) * 16 + ((t / A1) ^ t2d_swz<A1>(t % A1)); }, 16 * A1); // C1 writes
  for (int qq = 0; qq < A1; ++qq) audit([&](int t) { return (((t >> 4)) * A1 + qq) * 16 + ((t & 15) ^ t2d_swz<A1>(qq)); }, 16 * A1);  // C2 reads
  return worst;
}
extern "C" int emu_t2d_conflicts(int A1, int A2) {
  if (A1 == 8 && A2 == 8) return emu_t2d_conflicts_c<8, 8>();
  if (A1 == 16 && A2 == 8) return emu_t2d_conflicts_c<16, 8>();
  if (A1 == 8 && A2 == 16) return emu_t2d_conflicts_c<8, 16>();
  if (A1 == 16 && A2 == 16) return emu_t2d_conflicts_c<16, 16>();
  return -1;
}

// ---- cluster-fused tiled 2-D plan: the CL CTAs of one cluster stepped phase by phase (barriers = phase ends)
template <int A1, int A2, int CL, int SIGN>
static void emu_t2d_cluster_run(const float* in, float* out) {
  using namespace pf;
  using G = T2D<A1, A2>; using K = T2DC<A1, A2, CL>;
  std::vector<float> tab(2 * ((size_t)G::N1 + G::N2 + G::NC));
  t2d_fill_tables<float, A1, A2>(tab.data());
  const cf* twA = reinterpret_cast<const cf*>(tab.data());
  const cf* twC = twA + G::N2;
  const cf* tw2d = twC + G::N1;
  std::vector<std::vector<cf>> tile(CL, std::vector<cf>(K::TILE)), park(CL, std::vector<cf>((size_t)K::QC * K::PARK_BLOCK));
  cf* bases[CL];
  for (int c = 0; c < CL; ++c) bases[c] = park[c].data();
  const T2DClusterSink<float, A1, A2, CL, EmuRemote> sink{EmuRemote{bases}};
  const cf* x = reinterpret_cast<const cf*>(in);
  cf* X = reinterpret_cast<cf*>(out);
  for (int qa = 0; qa < K::QA; ++qa)
    for (int rank = 0; rank < CL; ++rank) {
      const int c = rank + CL * qa;
      for (int t = 0; t < G::TA; ++t) t2d_A1<A1, A2, SIGN, float>(t, x + 16 * c, twA, tile[rank].data());
      for (int t = 0; t < G::TA; ++t) t2d_A2<A1, A2, SIGN, float>(t, c, tile[rank].data(), tw2d, sink);
    }
  for (int qc = 0; qc < K::QC; ++qc)
    for (int rank = 0; rank < CL; ++rank) {
      const int d = rank + CL * qc;
      for (int t = 0; t < G::TC; ++t) t2d_C1<A1, A2, SIGN, float>(t, park[rank].data() + (size_t)qc * K::PARK_BLOCK, twC, tile[rank].data());
      for (int t = 0; t < G::TC; ++t) t2d_C2<A1, A2, SIGN, float>(t, tile[rank].data(), X + 16 * d);
    }
}
extern "C" int emu_t2d_cluster(int A1, int A2, int CL, int dir, const float* in, float* out) {
#define T2C(a1, a2, cl) if (A1 == a1 && A2 == a2 && CL == cl) { if (dir == 0) emu_t2d_cluster_run<a1, a2, cl, -1>(in, out); else emu_t2d_cluster_run<a1, a2, cl, +1>(in, out); return 0; }
  T2C(8, 8, 8) T2C(8, 8, 4) T2C(16, 8, 8) T2C(16, 16, 8) T2C(16, 16, 16)
#undef T2C
  return -1;
}

// ---- tiled 2-D plan, general radices (k_t2dg_A / k_t2dg_C): 256 threads per tile
template <int A1, int A2, int SIGN>
static void emu_t2dg_run(const float* in, float* out) {
  using namespace pf;
  using G = T2D<A1, A2>;
  std::vector<float> tab(2 * ((size_t)G::N1 + G::N2 + G::NC));
  t2d_fill_tables<float, A1, A2>(tab.data());
  const cf* twA = reinterpret_cast<const cf*>(tab.data());
  const cf* twC = twA + G::N2;
  const cf* tw2d = twC + G::N1;
  std::vector<cf> S(G::NC), tileA(16 * G::N2), tileC(16 * G::N1);
  const cf* x = reinterpret_cast<const cf*>(in);
  cf* X = reinterpret_cast<cf*>(out);
  for (int c = 0; c < G::N1 / 16; ++c) {
    for (int t = 0; t < G::TA; ++t) t2d_A1<A1, A2, SIGN, float>(t, x + 16 * c, twA, tileA.data());
    for (int t = 0; t < 256; ++t) t2dg_A2<A1, A2, SIGN, float>(t, c, tileA.data(), tw2d, T2DScratchSink<float, G::N1>{S.data()});
  }
  for (int d = 0; d < G::N2 / 16; ++d) {
    for (int t = 0; t < 256; ++t) t2dg_C1<A1, A2, SIGN, float>(t, S.data() + (size_t)16 * d * G::N1, twC, tileC.data());
    for (int t = 0; t < 256; ++t) t2dg_C2<A1, A2, SIGN, float>(t, tileC.data(), X + 16 * d);
  }
}
#define PF_T2DG_SHAPES(X) X(6, 5) X(6, 6) X(8, 6) X(10, 8) X(12, 8) X(12, 12) X(16, 10) X(16, 12) X(16, 15) X(8, 8) X(16, 8) X(16, 16)
extern "C" int emu_t2dg(int A1, int A2, int dir, const float* in, float* out) {
#define X(a1, a2) if (A1 == a1 && A2 == a2) { if (dir == 0) emu_t2dg_run<a1, a2, -1>(in, out); else emu_t2dg_run<a1, a2, +1>(in, out); return 0; }
  PF_T2DG_SHAPES(X)
#undef X
  return -1;
}
// conflict audit of the padded pass-C tile and the pass-A tile for general A (1 = conflict free)
template <int A1, int A2> static int emu_t2dg_conflicts_c() {
  int worst = 1;
  auto audit = [&](auto addr, auto active, int nthreads) {
    for (int h = 0; h < nthreads; h += 16) { int cnt[16] = {0}; for (int l = 0; l < 16; ++l) if (active(h + l)) cnt[addr(h + l) & 15]++; for (int b = 0; b < 16; ++b) if (cnt[b] > worst) worst = cnt[b]; }
  };
  auto all = [](int) { return true; };
  for (int ka = 0; ka < 16; ++ka) audit([&](int t) { return (ka * A2 + (t >> 4)) * 16 + (t & 15); }, all, 16 * A2);          // A1 writes
  for (int q = 0; q < A2; ++q) audit([&](int t) { return ((t >> 4) * A2 + q) * 16 + (t & 15); }, all, 256);                    // A2 reads
  for (int ka = 0; ka < 16; ++ka) audit([&](int t) { return (ka * A1 + (t & 15)) * 16 + ((t >> 4) ^ (t & 15)); }, [&](int t) { return (t & 15) < A1; }, 256);   // C1 writes
  for (int q = 0; q < A1; ++q) audit([&](int t) { return ((t >> 4) * A1 + q) * 16 + ((t & 15) ^ q); }, all, 256);              // C2 reads
  return worst;
}
extern "C" int emu_t2dg_conflicts(int A1, int A2) {
#define X(a1, a2) if (A1 == a1 && A2 == a2) return emu_t2dg_conflicts_c<a1, a2>();
  PF_T2DG_SHAPES(X)
#undef X
  return -1;
}

// ---- general tiled plan in double precision (the shapes the double launcher instantiates)
template <int A1, int A2, int SIGN>
static void emu_t2dg_run_d(const double* in, double* out) {
  using namespace pf;
  using G = T2D<A1, A2>;
  std::vector<double> tab(2 * ((size_t)G::N1 + G::N2 + G::NC));
  t2d_fill_tables<double, A1, A2>(tab.data());
  const cd* twA = reinterpret_cast<const cd*>(tab.data());
  const cd* twC = twA + G::N2;
  const cd* tw2d = twC + G::N1;
  std::vector<cd> S(G::NC), tileA(16 * G::N2), tileC(16 * G::N1);
  const cd* x = reinterpret_cast<const cd*>(in);
  cd* X = reinterpret_cast<cd*>(out);
  for (int c = 0; c < G::N1 / 16; ++c) {
    for (int t = 0; t < G::TA; ++t) t2d_A1<A1, A2, SIGN, double>(t, x + 16 * c, twA, tileA.data());
    for (int t = 0; t < 256; ++t) t2dg_A2<A1, A2, SIGN, double>(t, c, tileA.data(), tw2d, T2DScratchSink<double, G::N1>{S.data()});
  }
  for (int d = 0; d < G::N2 / 16; ++d) {
    for (int t = 0; t < 256; ++t) t2dg_C1<A1, A2, SIGN, double>(t, S.data() + (size_t)16 * d * G::N1, twC, tileC.data());
    for (int t = 0; t < 256; ++t) t2dg_C2<A1, A2, SIGN, double>(t, tileC.data(), X + 16 * d);
  }
}
extern "C" int emu_t2dg_double(int A1, int A2, int dir, const double* in, double* out) {
#define X(a1, a2) if (A1 == a1 && A2 == a2) { if (dir == 0) emu_t2dg_run_d<a1, a2, -1>(in, out); else emu_t2dg_run_d<a1, a2, +1>(in, out); return 0; }
  X(8, 8) X(16, 8) X(16, 16)
#undef X
  return -1;
}
