// Throughput of the SFU-class operations the fused kernels lean on (per SM, all four schedulers busy):
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 4) { unsigned u = __float_as_uint(v[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
      if (OP == 6) { unsigned short h; asm volatile("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(v[i])); v[i] = __uint_as_float((unsigned)h << 16); }
      if (OP == 7) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 8) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 9) asm volatile("max.f32 %0, %0, %0;" : "+f"(v[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int perop) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 512>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(out, cyc, 0.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double ops = 512.0 * 16 * ITER * perop;
  printf("%-28s %8.2f results / clk / SM   (%.0f clk)\n", name, ops / avg, avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.approx.f32", 1); run<1>("rcp.approx.f32", 1); run<2>("tanh.approx.f32", 1);
  run<3>("tanh.approx.bf16x2 (x2)", 2); run<4>("ex2.approx.bf16x2 (x2)", 2); run<7>("tanh.approx.f16x2 (x2)", 2);
  run<8>("rsqrt.approx.f32", 1); run<5>("fma.rn.f32", 1); run<6>("cvt.rn.bf16.f32", 1); run<9>("max.f32", 1);
  return 0;
}
