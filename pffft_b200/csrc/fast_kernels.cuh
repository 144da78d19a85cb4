// fast_kernels.cuh -- size-tuned sm_100a kernels for the headline sizes.
//
// c2c N=1024 fp32 (BASELINE configs C2/C5): ONE TRANSFORM PER WARP.
//   1024 = 32 x 32.  Lane n1 owns the 32 points x[n1 + 32*n2] (64 registers), runs a complete
//   radix-32 register FFT over n2, multiplies by W_1024^{n1*k2}, the warp transposes the 32x32
//   tile through a padded shared-memory tile (conflict-free 64-bit accesses, __syncwarp only --
//   no CTA barrier in the steady state), lane k2 runs a second radix-32 FFT over n1 and owns
//   X[k2 + 32*k1]: every global access of the warp is a fully coalesced 256-byte row.
//   HBM traffic: one 8 KiB read and one 8 KiB write per transform (the algorithmic minimum).
//   This replaces the reference's 7 sweeps (uninterleave, 4x passf4_ps, cplx_finalize,
//   zreorder; src/pffft_priv_impl.h:1465-1532).
//
//   Two feeding schemes share the arithmetic:
//     k_c1024_ldg  : lanes load their points straight into registers (LDG.64), latency hidden by
//                    occupancy (16 warps/SM).
//     k_c1024_bulk : each warp owns a 2-deep ring of 8 KiB shared-memory stages filled by the TMA
//                    engine with 1-D bulk-async copies (cp.async.bulk ... mbarrier::complete_tx,
//                    SASS UBLKCP) issued two transforms ahead; the consumed stage doubles as the
//                    transpose tile.  Loads cost no registers and no issue slots.
#pragma once
#include "butterfly.cuh"
#include "layout.cuh"

namespace pf {

constexpr int kW1024Tile = 32 * 33;   // padded 32x32 tile, in complex elements (8448 bytes)

PF_HD constexpr int brev5(int p) { return ct::bitrev(p, 5); }

// phase A: v[p] = x[lane + 32*brev5(p)] on entry.  Row FFT over n2, twiddle, write tile row `lane`.
//   tw[k2*32 + n1] = exp(-2 pi i n1 k2 / 1024)
template <int SIGN>
PF_HD void w1024_rows(cf (&v)[32], int lane, const cf* tw, cf* tile) {
  reg_fft<32, SIGN>(v);                       // v[k2] = sum_n2 x[lane + 32 n2] w32^(n2 k2)
  tile[lane * 33] = v[0];
#pragma unroll
  for (int k2 = 1; k2 < 32; ++k2) tile[lane * 33 + k2] = cmul_dir<SIGN>(v[k2], tw[k2 * 32 + lane]);
}
// phase B: gather column `lane` (= k2) in bit-reversed register order, column FFT over n1.
//   on exit v[k1] = X[lane + 32*k1]
template <int SIGN>
PF_HD void w1024_cols(cf (&v)[32], int lane, const cf* tile) {
#pragma unroll
  for (int p = 0; p < 32; ++p) v[p] = tile[brev5(p) * 33 + lane];
  reg_fft<32, SIGN>(v);
}

// (device helpers follow)
#ifdef __CUDACC__
PF_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PF_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
PF_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PF_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
PF_D void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
PF_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
PF_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

PF_D cf ld_stream(const cf* p) { float2 t = __ldcs(reinterpret_cast<const float2*>(p)); return mk<float>(t.x, t.y); }
PF_D void st_stream(cf* p, cf v) { __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }

// store the finished transform: lane owns X[lane + 32*k1].  ZOUT selects the reference's z-domain layout
// (pffft_transform): the z image is assembled in the warp's tile (4-byte scatters, granule-swizzled so they are
// bank-conflict free) and leaves with coalesced 128-bit stores.
template <bool ZOUT>
PF_D void w1024_store(const cf (&v)[32], int lane, cf* dst, cf* tile) {
  if (!ZOUT) {
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) st_stream(dst + lane + 32 * k1, v[k1]);
  } else {
    float* tf = reinterpret_cast<float*>(tile);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
      const int p = zpos_complex(lane + 32 * k1, 1024);
      tf[zswz(p)] = v[k1].x; tf[zswz(p + 4)] = v[k1].y;
    }
    __syncwarp();
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int e = 4 * (lane + 32 * j);                    // first float of granule lane + 32 j
      __stcs(d4 + lane + 32 * j, *reinterpret_cast<const float4*>(tf + zswz(e)));
    }
    __syncwarp();
  }
}
// first-phase loader; ZIN: the whole z-domain transform is copied into the tile with coalesced 128-bit loads
// (granule-swizzled), then every lane gathers its 32 points from there
template <bool ZIN>
PF_D void w1024_load(cf (&v)[32], int lane, const cf* src, cf* tile) {
  if (!ZIN) {
#pragma unroll
    for (int p = 0; p < 32; ++p) v[p] = ld_stream(src + lane + 32 * brev5(p));
  } else {
    float* tf = reinterpret_cast<float*>(tile);
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int e = 4 * (lane + 32 * j);
      *reinterpret_cast<float4*>(tf + zswz(e)) = __ldcs(s4 + lane + 32 * j);
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const int q = zpos_complex(lane + 32 * brev5(p), 1024);
      v[p] = mk<float>(tf[zswz(q)], tf[zswz(q + 4)]);
    }
    __syncwarp();                                           // the tile is rewritten by the row phase
  }
}

// ---------------------------------------------------------------- register-fed variant
template <int SIGN, int WARPS, int MINB, bool ZIN, bool ZOUT>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_c1024_ldg(const cf* __restrict__ in, cf* __restrict__ out, long long batch, const cf* __restrict__ tw_g) {
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cf* tw = reinterpret_cast<cf*>(pf_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf* tile = tw + 1024 + warp * kW1024Tile;
  for (int i = threadIdx.x; i < 1024; i += WARPS * 32) tw[i] = tw_g[i];
  __syncthreads();
  const long long stride = (long long)gridDim.x * WARPS;
  for (long long t = (long long)blockIdx.x * WARPS + warp; t < batch; t += stride) {
    const cf* src = in + t * 1024;
    cf v[32];
    w1024_load<ZIN>(v, lane, src, tile);
    w1024_rows<SIGN>(v, lane, tw, tile);
    __syncwarp();
    w1024_cols<SIGN>(v, lane, tile);
    __syncwarp();                               // tile is rewritten by the next iteration
    w1024_store<ZOUT>(v, lane, out + t * 1024, tile);
  }
}

// ---------------------------------------------------------------- TMA bulk-copy fed variant
// shared memory: [tw 8 KiB][per warp: 2 stages x 8448 B][mbarriers: WARPS x 2 x 8 B]
template <int SIGN, int WARPS, int MINB, bool ZOUT>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_c1024_bulk(const cf* __restrict__ in, cf* __restrict__ out, long long batch, const cf* __restrict__ tw_g) {
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cf* tw = reinterpret_cast<cf*>(pf_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf* stage0 = tw + 1024 + warp * (2 * kW1024Tile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tw + 1024 + WARPS * (2 * kW1024Tile)) + warp * 2;
  for (int i = threadIdx.x; i < 1024; i += WARPS * 32) tw[i] = tw_g[i];
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); fence_proxy_async(); }
  __syncthreads();

  const long long stride = (long long)gridDim.x * WARPS;
  const long long first = (long long)blockIdx.x * WARPS + warp;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const long long t = first + s * stride;
      if (t < batch) { mbar_expect_tx(&bars[s], 8192); bulk_g2s(stage0 + s * kW1024Tile, in + t * 1024, 8192, &bars[s]); }
    }
  }
  int it = 0;
  for (long long t = first; t < batch; t += stride, ++it) {
    const int s = it & 1;
    cf* buf = stage0 + s * kW1024Tile;
    mbar_wait(&bars[s], (it >> 1) & 1);
    cf v[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) v[p] = buf[lane + 32 * brev5(p)];
    __syncwarp();                               // every lane has its points: the stage becomes the tile
    w1024_rows<SIGN>(v, lane, tw, buf);
    __syncwarp();
    w1024_cols<SIGN>(v, lane, buf);
    __syncwarp();                               // tile consumed
    if (ZOUT) w1024_store<true>(v, lane, out + t * 1024, buf);   // the z image is assembled in the stage: store first
    if (lane == 0) {                            // hand the stage back to the TMA engine
      const long long tn = t + 2 * stride;
      if (tn < batch) {
        fence_proxy_async();                    // order our generic-proxy accesses before the async-proxy write
        mbar_expect_tx(&bars[s], 8192);
        bulk_g2s(buf, in + tn * 1024, 8192, &bars[s]);
      }
    }
    if (!ZOUT) w1024_store<false>(v, lane, out + t * 1024, buf);
  }
}
#endif  // __CUDACC__


// ---------------------------------------------------------------------------------------------
// Small complex sizes Nc = 32*R2 (R2 = 1,2,4,8 -> N = 32,64,128,256): the same warp machinery, with the
// warp's 1024-point chunk holding TW = 32/R2 whole transforms.  Lane l loads the chunk's elements l + 32 m
// (coalesced rows, exactly as for N=1024); m = j*R2 + n2 selects transform j and the stride-32 digit n2.
//   phase A: per transform j a radix-R2 register FFT over n2, * W_Nc^{l k2}, tile[l][j*R2 + k2]
//   phase B: lane c = (j, k2) runs the radix-32 FFT down column c  ->  X_j[k2 + R2*k1]
// tw[k2*32 + l] = exp(-2 pi i l k2 / Nc).  Canonical (ordered) layout in and out.
// ---------------------------------------------------------------------------------------------
template <int R2> PF_HD constexpr int brevR2(int p) { return ct::bitrev(p, ct::ilog2(R2)); }

template <int R2, int SIGN>
PF_HD void wsmall_rows(cf (&v)[32], int lane, const cf* tw, cf* tile) {
  // v[j*R2 + p] = chunk[lane + 32*(j*R2 + brevR2(p))] on entry
  if constexpr (R2 == 8) { dit_fft<8, SIGN, 0, 1>(v); dit_fft<8, SIGN, 8, 1>(v); dit_fft<8, SIGN, 16, 1>(v); dit_fft<8, SIGN, 24, 1>(v); }
  if constexpr (R2 == 4) {
    dit_fft<4, SIGN, 0, 1>(v); dit_fft<4, SIGN, 4, 1>(v); dit_fft<4, SIGN, 8, 1>(v); dit_fft<4, SIGN, 12, 1>(v);
    dit_fft<4, SIGN, 16, 1>(v); dit_fft<4, SIGN, 20, 1>(v); dit_fft<4, SIGN, 24, 1>(v); dit_fft<4, SIGN, 28, 1>(v);
  }
  if constexpr (R2 == 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { const cf a = v[2 * j], b = v[2 * j + 1]; v[2 * j] = a + b; v[2 * j + 1] = a - b; }
  }
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    const int k2 = m % R2;
    tile[lane * 33 + m] = (k2 == 0) ? v[m] : cmul_dir<SIGN>(v[m], tw[k2 * 32 + lane]);
  }
}

#ifdef __CUDACC__

#ifdef __CUDACC__
// z-domain images of the `nvalid` transforms of a warp chunk, staged in the warp tile (granule-swizzled per transform)
// and moved to / from global memory with coalesced 128-bit accesses.  NC complex points = 2*NC floats per transform.
template <int NC> PF_D void chunk_z_out(const float* tf, float* dst, int nvalid, int lane) {
  const int ngran = nvalid * (2 * NC) / 4;
  for (int G = lane; G < ngran; G += 32) {
    const int e = 4 * G, j = e / (2 * NC), p = e - j * (2 * NC);
    __stcs(reinterpret_cast<float4*>(dst) + G, *reinterpret_cast<const float4*>(tf + j * (2 * NC) + zswz(p)));
  }
}
template <int NC> PF_D void chunk_z_in(float* tf, const float* src, int nvalid, int lane) {
  const int ngran = nvalid * (2 * NC) / 4;
  for (int G = lane; G < ngran; G += 32) {
    const int e = 4 * G, j = e / (2 * NC), p = e - j * (2 * NC);
    *reinterpret_cast<float4*>(tf + j * (2 * NC) + zswz(p)) = __ldcs(reinterpret_cast<const float4*>(src) + G);
  }
}
#endif

// one warp = one 1024-point chunk = 32/R2 transforms; grid-stride over chunks.
// ZIN / ZOUT: the transform's input / output is in the reference's z-domain layout (pffft_transform), so the
// ordered and unordered entry points share one arithmetic path and stay bit-identical to each other.
template <int R2, int SIGN, int WARPS, int MINB, bool ZIN, bool ZOUT>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_warp_small(const cf* in, cf* out, long long batch, const cf* __restrict__ tw_g) {
  constexpr int NC = 32 * R2;
  constexpr int TW = 32 / R2;                   // transforms per warp chunk
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cf* tw = reinterpret_cast<cf*>(pf_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf* tile = tw + NC + warp * kW1024Tile;
  for (int i = threadIdx.x; i < NC; i += WARPS * 32) tw[i] = tw_g[i];
  __syncthreads();
  const long long nchunks = (batch + TW - 1) / TW;
  const long long stride = (long long)gridDim.x * WARPS;
  for (long long c = (long long)blockIdx.x * WARPS + warp; c < nchunks; c += stride) {
    const cf* src = in + c * 1024;
    cf* dst = out + c * 1024;
    const long long left = batch - c * TW;      // transforms of this chunk that exist (>= 1)
    const int nvalid = left >= TW ? TW : (int)left;
    cf v[32];
    float* tf = reinterpret_cast<float*>(tile);
    if (ZIN) { chunk_z_in<NC>(tf, reinterpret_cast<const float*>(src), nvalid, lane); __syncwarp(); }
#pragma unroll
    for (int j = 0; j < TW; ++j)
#pragma unroll
      for (int p = 0; p < R2; ++p) {
        const int n = lane + 32 * brevR2<R2>(p);                       // index inside transform j
        if (j >= nvalid) v[j * R2 + p] = mk<float>(0.f, 0.f);
        else if (!ZIN) v[j * R2 + p] = ld_stream(src + j * NC + n);
        else { const int q = zpos_complex(n, NC); v[j * R2 + p] = mk<float>(tf[j * 2 * NC + zswz(q)], tf[j * 2 * NC + zswz(q + 4)]); }
      }
    if (ZIN) __syncwarp();
    wsmall_rows<R2, SIGN>(v, lane, tw, tile);
    __syncwarp();
    w1024_cols<SIGN>(v, lane, tile);            // column `lane` = (j = lane / R2, k2 = lane % R2)
    __syncwarp();
    const int j = lane / R2, k2 = lane % R2;
    if (ZOUT) {
      if (j < nvalid) {
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) { const int q = zpos_complex(k2 + R2 * k1, NC); tf[j * 2 * NC + zswz(q)] = v[k1].x; tf[j * 2 * NC + zswz(q + 4)] = v[k1].y; }
      }
      __syncwarp();
      chunk_z_out<NC>(tf, reinterpret_cast<float*>(dst), nvalid, lane);
      __syncwarp();
    } else if (R2 <= 2) {
      // 8/16-byte pieces per lane would scatter: go back through the tile and store whole 256-byte rows
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) { const int e = j * NC + k2 + R2 * k1; tile[(e >> 5) * 33 + (e & 31)] = v[k1]; }
      __syncwarp();
      const int rows = nvalid * R2;             // 32-element rows of the chunk that hold existing transforms
#pragma unroll
      for (int r = 0; r < 32; ++r) if (r < rows) st_stream(dst + 32 * r + lane, tile[r * 33 + lane]);
      __syncwarp();
    } else if (j < nvalid) {
      cf* d = dst + j * NC + k2;
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) st_stream(d + R2 * k1, v[k1]);
    }
  }
}
#endif


// ---------------------------------------------------------------------------------------------
// Non-power-of-two complex sizes N = 32*R2, R2 in {3,5,6,9,10,12,15} (N = 96,160,192,288,320,384,480): the warp
// machinery again, with a mixed-radix register DFT (radix 3/5 butterflies) as the row transform.  A warp chunk holds
// TW = floor(32/R2) transforms; lane l loads x_j[l + 32*n2]; phase A: TW DFTs of size R2 per lane, * W_N^{l k2},
// tile[l][j*R2 + k2]; phase B: lanes c = (j,k2) < TW*R2 run the radix-32 column FFT -> X_j[k2 + R2*k1]; the result
// goes back through the tile so that global stores are whole 256-byte rows.  tw[k2*32 + l] = exp(-2 pi i l k2 / N).
// ---------------------------------------------------------------------------------------------
template <int R2, int SIGN>
PF_HD void wmixed_rows(cf (&v)[32], int lane, const cf* tw, cf* tile) {
  constexpr int TW = 32 / R2;
#pragma unroll
  for (int j = 0; j < TW; ++j) dft_small<R2, SIGN>(&v[j * R2]);
#pragma unroll
  for (int m = 0; m < TW * R2; ++m) {
    const int k2 = m % R2;
    tile[lane * 33 + m] = (k2 == 0) ? v[m] : cmul_dir<SIGN>(v[m], tw[k2 * 32 + lane]);
  }
}

// real transforms on the same kernel (REAL): N = 64*R2 real points are NC = 32*R2 packed complex points
// (z[i] = x[2i] + i x[2i+1]); the rotation that turns the packed spectrum into the real one (and back) is applied
// while the chunk sits in the warp tile:  X[k] = ((s.x + u.y), (s.y - u.x))/2,  s = Z[k] + conj Z[NC-k],
// u = W_N^k (Z[k] - conj Z[NC-k]);   Z'[i] = (s.x - u.y, s.y + u.x),  s = X[i] + conj X[NC-i], u = conj(W_N^i)(X[i] - conj X[NC-i]).
template <int NC> PF_HD int chunk_tile_idx(int j, int k) { const int e = j * NC + k; return (e >> 5) * 33 + (e & 31); }

#ifdef __CUDACC__
template <int R2, int SIGN, int WARPS, int MINB, bool ZIN, bool ZOUT, bool REAL>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_warp_mixed(const cf* in, cf* out, long long batch, const cf* __restrict__ tw_g, const cf* __restrict__ twr, int grp) {
  // grp > 1 (complex canonical input only): transform g reads the decimated sub-sequence
  //   in[(g / grp) * grp * NC + (g % grp) + n * grp]   -- the rows of a two-pass (split) plan
  constexpr int NC = 32 * R2;
  constexpr int TW = 32 / R2;                   // transforms per warp chunk
  constexpr int COLS = TW * R2;                 // active lanes in phase B
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cf* tw = reinterpret_cast<cf*>(pf_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf* tile = tw + NC + warp * kW1024Tile;
  for (int i = threadIdx.x; i < NC; i += WARPS * 32) tw[i] = tw_g[i];
  __syncthreads();
  const long long nchunks = (batch + TW - 1) / TW;
  const long long stride = (long long)gridDim.x * WARPS;
  for (long long c = (long long)blockIdx.x * WARPS + warp; c < nchunks; c += stride) {
    const cf* src = in + c * (TW * NC);
    cf* dst = out + c * (TW * NC);
    const long long left = batch - c * TW;
    const int nvalid = left >= TW ? TW : (int)left;
    const int rows = nvalid * R2;               // 32-element rows of the chunk that hold existing transforms
    cf v[32];
    float* tf = reinterpret_cast<float*>(tile);
#pragma unroll
    for (int m = 0; m < 32; ++m) v[m] = mk<float>(0.f, 0.f);
    if (REAL && SIGN > 0) {
      // ---- backward real: spectrum (canonical or z-domain) -> tile -> packed half-length spectrum Z'
      if (ZIN) chunk_z_in<NC>(tf, reinterpret_cast<const float*>(src), nvalid, lane);
      else {
#pragma unroll
        for (int r = 0; r < COLS; ++r) if (r < rows) tile[r * 33 + lane] = ld_stream(src + 32 * r + lane);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
          if (j >= nvalid) continue;
          const int i = lane + 32 * n2;
          auto X = [&](int k) -> cf {
            if (ZIN) { const int q = zpos_real(k, 2 * NC); return mk<float>(tf[j * 2 * NC + zswz(q)], tf[j * 2 * NC + zswz(q + 4)]); }
            return tile[chunk_tile_idx<NC>(j, k)];
          };
          if (i == 0) { const cf s0 = X(0); v[j * R2 + n2] = mk<float>(s0.x + s0.y, s0.x - s0.y); }
          else {
            const cf a = X(i), b = conj(X(NC - i));
            const cf s = a + b, d = a - b;
            const float2 wv = __ldg(reinterpret_cast<const float2*>(twr + i));
            const cf u = cmul_dir<+1>(d, mk<float>(wv.x, wv.y));
            v[j * R2 + n2] = mk<float>(s.x - u.y, s.y + u.x);
          }
        }
      __syncwarp();
    } else {
      if (ZIN) { chunk_z_in<NC>(tf, reinterpret_cast<const float*>(src), nvalid, lane); __syncwarp(); }
#pragma unroll
      for (int j = 0; j < TW; ++j)
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
          const int n = lane + 32 * n2;
          if (j < nvalid) {
            if (ZIN) { const int q = zpos_complex(n, NC); v[j * R2 + n2] = mk<float>(tf[j * 2 * NC + zswz(q)], tf[j * 2 * NC + zswz(q + 4)]); }
            else if (grp == 1) v[j * R2 + n2] = ld_stream(src + j * NC + n);
            else {
              const long long g = c * TW + j, t = g / grp;
              const int n1 = (int)(g - t * grp);
              const float2 w2 = __ldg(reinterpret_cast<const float2*>(in + t * (long long)grp * NC + n1 + (long long)n * grp));
              v[j * R2 + n2] = mk<float>(w2.x, w2.y);
            }
          }
        }
      if (ZIN) __syncwarp();
    }
    wmixed_rows<R2, SIGN>(v, lane, tw, tile);
    __syncwarp();
    w1024_cols<SIGN>(v, lane, tile);            // lanes >= COLS transform unused columns (harmless, never stored)
    __syncwarp();
    const int j = lane / R2, k2 = lane % R2;
    const bool mine = lane < COLS && j < nvalid;
    if (REAL && SIGN < 0) {
      // ---- forward real: packed spectrum Z -> tile (natural order) -> X, canonical or z-domain
      if (mine) {
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) tile[chunk_tile_idx<NC>(j, k2 + R2 * k1)] = v[k1];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < COLS; ++r) {
        if (r >= rows) continue;
        const int e = 32 * r + lane, jj = e / NC, k = e - jj * NC;
        cf x;
        if (k == 0) { const cf z0 = tile[chunk_tile_idx<NC>(jj, 0)]; x = mk<float>(z0.x + z0.y, z0.x - z0.y); }
        else {
          const cf a = tile[chunk_tile_idx<NC>(jj, k)], b = conj(tile[chunk_tile_idx<NC>(jj, NC - k)]);
          const float2 wv = __ldg(reinterpret_cast<const float2*>(twr + k));
          const cf s = a + b, d = a - b, u = cmul(d, mk<float>(wv.x, wv.y));
          x = mk<float>(0.5f * (s.x + u.y), 0.5f * (s.y - u.x));
        }
        if (!ZOUT) st_stream(dst + e, x); else v[r] = x;
      }
      if (ZOUT) {
        __syncwarp();                           // every lane has read its Z values: the tile becomes the z image
#pragma unroll
        for (int r = 0; r < COLS; ++r) {
          if (r >= rows) continue;
          const int e = 32 * r + lane, jj = e / NC, k = e - jj * NC;
          const int q = zpos_real(k, 2 * NC);
          tf[jj * 2 * NC + zswz(q)] = v[r].x; tf[jj * 2 * NC + zswz(q + 4)] = v[r].y;
        }
        __syncwarp();
        chunk_z_out<NC>(tf, reinterpret_cast<float*>(dst), nvalid, lane);
      }
      __syncwarp();
    } else if (ZOUT) {
      if (mine) {
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) { const int q = zpos_complex(k2 + R2 * k1, NC); tf[j * 2 * NC + zswz(q)] = v[k1].x; tf[j * 2 * NC + zswz(q + 4)] = v[k1].y; }
      }
      __syncwarp();
      chunk_z_out<NC>(tf, reinterpret_cast<float*>(dst), nvalid, lane);
      __syncwarp();
    } else {
      if (mine) {
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) tile[chunk_tile_idx<NC>(j, k2 + R2 * k1)] = v[k1];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < COLS; ++r) if (r < rows) st_stream(dst + 32 * r + lane, tile[r * 33 + lane]);
      __syncwarp();
    }
  }
}
#endif

}  // namespace pf
