/* A C99 host with no Python and no torch: cudaMalloc buffers, one dk_gemm (bias + GELU-erf epilogue) through the C ABI
 * of include/dkb200.h, result checked against a double-precision evaluation on the CPU.
 * Build: gcc -std=c99 abi_gemm.c -I include -I $CUDA/include -L diffusionkit_b200 -ldkb200 -L $CUDA/lib64 -lcudart -lm */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dkb200.h"

static uint16_t f2bf(float f) { /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static float frand(uint32_t* s) {
  *s = *s * 1664525u + 1013904223u;
  return ((float)((*s >> 8) & 0xFFFF) / 65536.0f) * 2.0f - 1.0f;
}

int main(void) {
  const int M = 300, N = 264, K = 200; /* ragged in every dimension */
  dk_ctx* ctx = NULL;
  if (dk_ctx_create(0, &ctx) != 0) {
    printf("FAIL ctx: %s\n", dk_last_error());
    return 1;
  }
  uint16_t *hA = malloc(2u * M * K), *hW = malloc(2u * N * K), *hb = malloc(2u * N), *hO = malloc(2u * M * N);
  uint32_t seed = 12345u;
  for (int i = 0; i < M * K; ++i) hA[i] = f2bf(frand(&seed));
  for (int i = 0; i < N * K; ++i) hW[i] = f2bf(frand(&seed) * 0.07f);
  for (int i = 0; i < N; ++i) hb[i] = f2bf(frand(&seed) * 0.5f);
  void *dA, *dW, *db, *dO;
  cudaMalloc(&dA, 2u * M * K);
  cudaMalloc(&dW, 2u * N * K);
  cudaMalloc(&db, 2u * N);
  cudaMalloc(&dO, 2u * M * N);
  cudaMemcpy(dA, hA, 2u * M * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW, 2u * N * K, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb, 2u * N, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, 2u * M * N);

  dk_gemm_args a;
  memset(&a, 0, sizeof a);
  a.dtype = DK_BF16;
  a.M = M;
  a.N = N;
  a.K = K;
  a.A = dA;
  a.lda = K;
  a.W = dW;
  a.ldw = K;
  a.out = dO;
  a.ldc = N;
  a.bias = db;
  a.act = DK_ACT_GELU_ERF;
  if (dk_gemm(ctx, &a, NULL) != 0) { /* NULL = the default stream */
    printf("FAIL gemm: %s\n", dk_last_error());
    return 1;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("FAIL sync\n");
    return 1;
  }
  cudaMemcpy(hO, dO, 2u * M * N, cudaMemcpyDeviceToHost);
  double num = 0.0, den = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = bf2f(hb[n]);
      for (int k = 0; k < K; ++k) acc += (double)bf2f(hA[m * K + k]) * (double)bf2f(hW[n * K + k]);
      const double ref = 0.5 * acc * (1.0 + erf(acc * 0.70710678118654752));
      const double d = (double)bf2f(hO[m * N + n]) - ref;
      num += d * d;
      den += ref * ref;
    }
  const double rel = sqrt(num / (den + 1e-30));
  /* an invalid call must be refused with a message, not crash */
  a.K = 12;
  const int rc_bad = dk_gemm(ctx, &a, NULL);
  printf("%s rel_l2=%.3e launches=%lld bad_rc=%d bad_msg=%s\n", (rel <= 4e-3 && rc_bad != 0) ? "OK" : "FAIL", rel,
         dk_ctx_launch_count(ctx), rc_bad, dk_last_error());
  dk_ctx_destroy(ctx);
  return (rel <= 4e-3 && rc_bad != 0) ? 0 : 1;
}
