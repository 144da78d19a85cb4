// layout.cuh -- index maps between the canonical ("ordered") spectrum layout and the reference's
// 4-lane internal ("unordered", z-domain) layout.
//
// The maps restate what pffft_zreorder does in the reference (src/pffft_priv_impl.h:1158-1193,
// reversed_copy :1125-1139, unreversed_copy :1141-1156) as closed-form index functions
// (SURVEY.md App. A), so a kernel can read or write the z-domain directly.  Units are scalar
// elements (floats or doubles); the imaginary part always lives 4 elements after the real part.
#pragma once
#include "common.cuh"

namespace pf {

// complex transform of N points: canonical complex index c in [0,N) -> element index of its real part.
// Memory is blocks of 8 four-lane vectors [r0 i0 r1 i1 r2 i2 r3 i3]; element u = c mod N/4 sits in
// lane u%4 of block u/4, quarter q = c / (N/4) selects the (r_q, i_q) vector pair.
PF_HD int zpos_complex(int c, int N) {
  const int nq = N >> 2;
  const int q = c / nq;
  const int u = c - q * nq;
  return 32 * (u >> 2) + 8 * q + (u & 3);
}

// real transform of N points: canonical slot k in [0,N/2) (slot 0 = (DC, Nyquist)) -> element index
// of its first component.  Quarter 0 holds bins [0,N/8), quarter 2 bins [N/4,3N/8) ascending;
// quarters 1 and 3 hold the remaining bins descending, with bins N/8 and 3N/8 parked in element 0.
PF_HD int zpos_real(int k, int N) {
  const int n8 = N >> 3;
  int q, u;
  if (k < n8)            { q = 0; u = k; }
  else if (k == n8)      { q = 1; u = 0; }
  else if (k < 2 * n8)   { q = 1; u = 2 * n8 - k; }
  else if (k < 3 * n8)   { q = 2; u = k - 2 * n8; }
  else if (k == 3 * n8)  { q = 3; u = 0; }
  else                   { q = 3; u = 4 * n8 - k; }
  return 32 * (u >> 2) + 8 * q + (u & 3);
}

// z-domain positions inside a SHARED-MEMORY staging copy of a spectrum are XOR-swizzled at 16-byte-granule level
// (element-index bits 2..4 ^= bits 5..7): consecutive bins land in 4-element groups 32 elements apart, which without
// the swizzle all hit the same 4 banks.  Granules stay intact, so the global side is copied with plain 128-bit accesses.
PF_HD int zswz(int p) { return p ^ (((p >> 5) & 7) << 2); }

template <bool REAL> PF_HD int zpos(int c, int N) { return REAL ? zpos_real(c, N) : zpos_complex(c, N); }

}  // namespace pf
