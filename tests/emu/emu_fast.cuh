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
