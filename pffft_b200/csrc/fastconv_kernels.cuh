// fastconv_kernels.cuh -- the whole overlap-save block in ONE kernel (BASELINE config C4).
//
// The reference processes a block with five sweeps over memory: memcpy -> pffft_transform(FORWARD) ->
// pffft_zconvolve_no_accu(Hf, 1/Nfft) -> pffft_transform(BACKWARD) -> memcpy (src/pffastconv.c:231-255).
// Here one CTA owns one block end to end: it reads its (overlapping, zero-padded) input window from the
// stream, runs the forward real FFT (N/2-point complex core of cta_kernels.cuh), applies
//     forward post-rotation  ->  x H[k] x 1/Nfft  ->  backward pre-rotation
// to each bin pair (k, Nc-k) in shared memory, runs the inverse core and stores only the valid
// `Nfft - filterLen + 1` samples.  HBM traffic per block = its input window + its valid outputs
// (8 B per output sample, the algorithmic minimum of SURVEY 8d, plus the filter-length overlap which
// neighbouring blocks share through L2); the filter spectrum stays L2-resident.
#pragma once
#include "cta_kernels.cuh"

namespace pf {

// One bin pair of   Z (packed half-length spectrum of the block)  ->  Z' (packed spectrum of block (*) h)
//   X[k]  = ((s.x + u.y), (s.y - u.x))/2,  X[M-k] = ((s.x - u.y), -(s.y + u.x))/2     with s,u as in real_post_pair
//   Y     = X * H * scale                                                   (H canonical: slot 0 = (H[0], H[M]))
//   Z'[k] = (S.x - U.y, S.y + U.x),       Z'[M-k] = (S.x + U.y, U.x - S.y)            S = Yk + conj Ym, U = (Yk - conj Ym) conj(W^k)
template <typename T>
PF_HD void fastconv_pair(cpx<T>* z, int k, int Nc, const cpx<T>* twr, const cpx<T>* Hc, T scale) {
  if (k == 0) {
    const cpx<T> z0 = z[0], zm = z[Nc / 2];
    const cpx<T> h0 = ldtab(Hc), hm = ldtab(Hc + Nc / 2);
    const T y0 = (z0.x + z0.y) * h0.x * scale;                    // DC and Nyquist are independent reals
    const T yM = (z0.x - z0.y) * h0.y * scale;                    // (ref pffft_priv_impl.h:1680-1683)
    z[0] = mk<T>(y0 + yM, y0 - yM);
    const cpx<T> ym = scale2(cmul(mk<T>(zm.x, -zm.y), hm), scale); // X[M/2] = conj Z[M/2]
    z[Nc / 2] = mk<T>(T(2) * ym.x, T(-2) * ym.y);                  // Z'[M/2] = 2 conj Y
    return;
  }
  const cpx<T> a = z[k], b = conj(z[Nc - k]);
  const cpx<T> w = ldtab(twr + k);
  const cpx<T> s = a + b, d = a - b;
  const cpx<T> u = cmul(d, w);
  const cpx<T> xk = scale2(s + mul_mi(u), T(0.5));
  const cpx<T> xm = emul(s - mul_mi(u), mk<T>(T(0.5), T(-0.5)));
  const cpx<T> yk = scale2(cmul(xk, ldtab(Hc + k)), scale);
  const cpx<T> ym = scale2(cmul(xm, ldtab(Hc + Nc - k)), scale);
  const cpx<T> S = yk + conj(ym), D = yk - conj(ym);
  const cpx<T> U = cmul_dir<+1>(D, w);                            // D * conj(W^k)
  z[k] = mk<T>(S.x - U.y, S.y + U.x);
  z[Nc - k] = mk<T>(S.x + U.y, U.x - S.y);
}

// inverse pass 1 fed from the natural-order shared buffer instead of global memory
template <int C, int SIGN, typename T>
PF_HD void k2_pass1_smem(int m, const cpx<T>* nat, const cpx<T>* tw1, cpx<T>* tile) {
  using K = K2<C>;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) v[p] = nat[m + K::BC * brev4(p)];
  reg_fft<16, SIGN>(v);
  const int jb = m / C, jc = m % C;
  tile[K::idx(0, jb, jc)] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[K::idx(ka, jb, jc)] = cmul_dir<SIGN>(v[ka], ldtab(tw1 + ka * K::BC + m));
}

struct FastconvParams {
  const float* x;           // input stream (device)
  float* y;                 // output stream (device)
  long long input_len;      // readable samples from x (zero padding beyond, ref pffastconv.c:231-233)
  long long n_full;         // blocks with a full Nfft window
  int stride;               // outputs (= input advance) per full block
  int tail_out;             // outputs of the one partial block that follows the full ones (0: none)
  float scale;              // 1/Nfft
  const cpx<float>* twr;    // exp(-2 pi i k / Nfft), k < Nfft/2
  const cpx<float>* Hc;     // filter spectrum, canonical layout
  const cpx<float>* tw1;    // tables of the 16x16xC core
  const cpx<float>* tw2;
};

// forward pass 1 for one PLANE of an interleaved complex stream (ES = 2: sample n of the plane is base[2 n]): the two-FFT
// complex mode of the reference (real and imaginary part convolved separately, src/pffastconv.c:212-247) without the
// de-interleaving copies -- the plane is picked apart by the loads and woven back by the stores
template <int C, typename T>
PF_HD void fastconv_pass1_plane(int m, const T* base, long long avail, const cpx<T>* tw1, cpx<T>* tile) {
  using K = K2<C>;
  cpx<T> v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const long long e = 2LL * (m + K::BC * brev4(p));               // first sample of pair i
    v[p] = mk<T>(e < avail ? base[2 * e] : T(0), e + 1 < avail ? base[2 * e + 2] : T(0));
  }
  reg_fft<16, -1>(v);
  const int jb = m / C, jc = m % C;
  tile[K::idx(0, jb, jc)] = v[0];
#pragma unroll
  for (int ka = 1; ka < 16; ++ka) tile[K::idx(ka, jb, jc)] = cmul_dir<-1>(v[ka], ldtab(tw1 + ka * K::BC + m));
}

#ifdef __CUDACC__
// ES = 1: real stream.  ES = 2: interleaved complex stream, work item = (block, plane); x/y point at the complex samples,
// input_len / stride / tail_out count complex samples.
template <int C, int MINB, int ES = 1>
__global__ void __launch_bounds__(16 * C, MINB) k_fastconv_fused(const FastconvParams p) {
  using K = K2<C>;
  constexpr int Nfft = 2 * K::NC;
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  cpx<float>* tile = reinterpret_cast<cpx<float>*>(pf_smem_raw);
  cpx<float>* nat = tile + K::NC;
  const int t = threadIdx.x;
  const long long nblk = p.n_full + (p.tail_out > 0 ? 1 : 0);
  const cpx<float>* tw1 = p.tw1;
  const cpx<float>* tw2 = p.tw2;
  const cpx<float>* twr = p.twr;
  const cpx<float>* Hc = p.Hc;
  for (long long w = blockIdx.x; w < nblk * ES; w += gridDim.x) {
    asm volatile("" : "+l"(tw1), "+l"(tw2), "+l"(twr), "+l"(Hc));   // keep table reads in the loop (see cta_kernels.cuh)
    const long long b = (ES == 1) ? w : (w >> 1);
    const int plane = (ES == 1) ? 0 : (int)(w & 1);                 // neighbouring CTAs take the two planes of one block
    const long long off = b * p.stride;
    const float* ibase = p.x + ES * off + plane;
    float* obase = p.y + ES * off + plane;
    const long long avail = p.input_len - off;
    const int out_count = (b < p.n_full) ? p.stride : p.tail_out;
    // ---- forward real FFT of the window
    if (ES == 2) fastconv_pass1_plane<C, float>(t, ibase, avail, tw1, tile);
    else {
    const bool vin = vec_aligned<float>(ibase);
    if (vin && avail >= (long long)Nfft) k2_pass1<C, L_R_TIME, -1, true, float>(t, ibase, Nfft, twr, avail, true, tw1, tile);
    else k2_pass1<C, L_R_TIME, -1, false, float>(t, ibase, Nfft, twr, avail, vin, tw1, tile);
    }
    __syncthreads();
    k2_pass2<C, -1, float>(t, tw2, tile);
    __syncthreads();
    cpx<float> u[16];
    k2_pass3<C, -1, float>(t, tile, u);
#pragma unroll
    for (int r = 0; r < 16 / C; ++r)
#pragma unroll
      for (int kc = 0; kc < C; ++kc) nat[k2_out_index<C>(t, r, kc)] = u[r * C + kc];
    __syncthreads();
    // ---- spectrum: rotate, multiply by the filter, rotate back (pairs k, Nc-k in place)
#pragma unroll 2
    for (int j = 0; j < 8; ++j) fastconv_pair<float>(nat, t + K::T * j, K::NC, twr, Hc, p.scale);
    __syncthreads();
    // ---- inverse
    k2_pass1_smem<C, +1, float>(t, nat, tw1, tile);
    __syncthreads();
    k2_pass2<C, +1, float>(t, tw2, tile);
    __syncthreads();
    k2_pass3<C, +1, float>(t, tile, u);
    const bool vok = vec_aligned<float>(obase);
#pragma unroll
    for (int r = 0; r < 16 / C; ++r)
#pragma unroll
      for (int kc = 0; kc < C; ++kc) {
        if (ES == 2) {
          const int e = 2 * k2_out_index<C>(t, r, kc);
          if (e < out_count) obase[2 * e] = u[r * C + kc].x;
          if (e + 1 < out_count) obase[2 * e + 2] = u[r * C + kc].y;
        } else store_elem<S_R_TIME, float>(obase, k2_out_index<C>(t, r, kc), u[r * C + kc], Nfft, out_count, vok);
      }
    // some threads may still be reading `tile` in pass 3 while others start the next block's pass 1
    __syncthreads();
  }
}
#endif  // __CUDACC__

}  // namespace pf
