// scratch: which access pattern streams HBM fastest on B200 (informs the elementwise kernels)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__global__ void copy_gridstride(const uint4* __restrict__ a, uint4* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] = a[i];
}
template <int U>
__global__ void copy_unroll(const uint4* __restrict__ a, uint4* __restrict__ b, long long n) {
  long long base = ((long long)blockIdx.x * blockDim.x) * U + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; ++k) { long long i = base + (long long)k * blockDim.x; if (i < n) v[k] = a[i]; }
#pragma unroll
  for (int k = 0; k < U; ++k) { long long i = base + (long long)k * blockDim.x; if (i < n) b[i] = v[k]; }
}
template <int U>
__global__ void copy_unroll_cs(const uint4* __restrict__ a, uint4* __restrict__ b, long long n) {
  long long base = ((long long)blockIdx.x * blockDim.x) * U + threadIdx.x;
  uint4 v[U];
#pragma unroll
  for (int k = 0; k < U; ++k) { long long i = base + (long long)k * blockDim.x; if (i < n) v[k] = __ldcs(a + i); }
#pragma unroll
  for (int k = 0; k < U; ++k) { long long i = base + (long long)k * blockDim.x; if (i < n) __stcs(b + i, v[k]); }
}
template <int U>
__global__ void copy_gs_unroll(const uint4* __restrict__ a, uint4* __restrict__ b, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = a[i + k * stride];
#pragma unroll
    for (int k = 0; k < U; ++k) b[i + k * stride] = v[k];
  }
  for (; i < n; i += stride) b[i] = a[i];
}
int main() {
  const long long bytes = 512LL << 20; const long long n = bytes / 16;
  uint4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMemset(a, 1, bytes);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(s); for (int i = 0; i < 10; ++i) launch(); cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e); printf("%-28s %.0f GB/s (r+w)\n", name, 2.0 * bytes * 10 / ms / 1e6);
  };
  run("gridstride 2368x256", [&] { copy_gridstride<<<2368, 256>>>(a, b, n); });
  run("gridstride 148*32x256", [&] { copy_gridstride<<<148 * 32, 256>>>(a, b, n); });
  run("oneshot U1", [&] { copy_unroll<1><<<(n + 255) / 256, 256>>>(a, b, n); });
  run("oneshot U4", [&] { copy_unroll<4><<<(n + 1023) / 1024, 256>>>(a, b, n); });
  run("oneshot U8", [&] { copy_unroll<8><<<(n + 2047) / 2048, 256>>>(a, b, n); });
  run("oneshot U4 cs", [&] { copy_unroll_cs<4><<<(n + 1023) / 1024, 256>>>(a, b, n); });
  run("gs-unroll4 2368x256", [&] { copy_gs_unroll<4><<<2368, 256>>>(a, b, n); });
  run("gs-unroll4 148*8x256", [&] { copy_gs_unroll<4><<<148 * 8, 256>>>(a, b, n); });
  run("gs-unroll8 148*4x512", [&] { copy_gs_unroll<8><<<148 * 4, 512>>>(a, b, n); });
  cudaMemcpy(b, a, bytes, cudaMemcpyDeviceToDevice);
  run("cudaMemcpy d2d", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
  return 0;
}
