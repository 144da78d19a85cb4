// ubench_fp32.cu -- issue-rate probe for the fp32 pipes of sm_100a: scalar FADD/FFMA against the packed
// FADD2/FMUL2/FFMA2 forms (add/mul/fma.f32x2), alone and mixed with integer work.  Not part of the product:
// the numbers decide how butterfly.cuh is written (profiles/r02_ubench_fp32.md).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench_fp32 tools/ubench_fp32.cu
#include <cuda_runtime.h>
#include <stdio.h>

#define ITER 4096
#define NACC 8

template <int MODE> __global__ void __launch_bounds__(256) k(float2* out, float2 seed, int n) {
  float2 a[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
  const float2 c = make_float2(seed.x * 0.5f, seed.y * 0.25f);
  const float2 d = make_float2(seed.y, seed.x);
  int q = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }                       // 2 FADD
      if (MODE == 1) { a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y); }    // 2 FFMA
      if (MODE == 2) { a[i] = __fadd2_rn(a[i], c); }                                          // 1 FADD2
      if (MODE == 3) { a[i] = __ffma2_rn(a[i], c, d); }                                       // 1 FFMA2
      if (MODE == 4) { a[i] = __fmul2_rn(a[i], c); }                                          // 1 FMUL2
      if (MODE == 5) { a[i] = __ffma2_rn(a[i], c, d); q = (q ^ (q >> 3)) + i; }               // FFMA2 + 2 alu
      if (MODE == 6) { a[i].x = fmaf(a[i].x, c.x, d.x); a[i].y = fmaf(a[i].y, c.y, d.y); q = (q ^ (q >> 3)) + i; }
      if (MODE == 7) { a[i].x = a[i].x + c.x; a[i].y = fmaf(a[i].y, c.y, d.y); }              // FADD + FFMA
      if (MODE == 8) { a[i] = __fadd2_rn(a[i], make_float2(-a[(i + 1) % NACC].x, -a[(i + 1) % NACC].y)); }  // a - b, packed (neg modifier?)
      if (MODE == 9) { const float2 s = make_float2(a[(i + 1) % NACC].y, a[(i + 1) % NACC].x); a[i] = __ffma2_rn(s, c, a[i]); }  // swapped operand
    }
  }
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NACC; ++i) { r.x += a[i].x; r.y += a[i].y; }
  r.x += q;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, int fp_per_iter, float2* out) {
  int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int grid = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, make_float2(1.0001f, 0.9999f), ITER);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, make_float2(1.0001f, 0.9999f), ITER);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double warp_instr = (double)grid * 8 * ITER * NACC * fp_per_iter;      // fp warp-instructions
  const double cyc = ms * 1e-3 * khz * 1e3;
  printf("%-28s %8.3f ms  fp warp-instr/clk/SM %6.3f  (per SMSP %5.3f)  clock %d MHz\n", name, ms,
         warp_instr / cyc / sms, warp_instr / cyc / sms / 4, khz / 1000);
}

int main() {
  float2* out; cudaMalloc(&out, sizeof(float2) * 148 * 8 * 256 * 2);
  run<0>("2 x FADD", 2, out);
  run<1>("2 x FFMA", 2, out);
  run<2>("FADD2", 1, out);
  run<3>("FFMA2", 1, out);
  run<4>("FMUL2", 1, out);
  run<5>("FFMA2 + 2 int", 1, out);
  run<6>("2 x FFMA + 2 int", 2, out);
  run<7>("FADD + FFMA", 2, out);
  run<8>("FADD2 with negated operand", 1, out);
  run<9>("FFMA2 with swapped operand", 1, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
