// Host-side internals shared by the translation units of libdkb200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dkb200.h"

typedef CUresult (*dk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct dk_ctx {
  int device;
  int sm_count;
  dk_encode_tiled_fn encode_tiled;
  long long launches;
  void* nccl_comm;  // ncclComm_t, lazily created by dk_comm_init
  void* nccl_lib;   // dlopen handle
};

void dk_set_error(const char* fmt, ...);

// Every entry point runs on ITS context's device whatever the calling thread's current device is (two pipelines on
// two GPUs in one process), and leaves the thread's device as it found it.
struct DkDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DkDeviceGuard(const dk_ctx* c) {
    if (c != nullptr && cudaGetDevice(&prev) == cudaSuccess && prev != c->device)
      switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~DkDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DkDeviceGuard(const DkDeviceGuard&) = delete;
  DkDeviceGuard& operator=(const DkDeviceGuard&) = delete;
};

#define DK_CHECK_CUDA(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      dk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));          \
      return -2;                                                                                   \
    }                                                                                              \
  } while (0)

#define DK_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      dk_set_error(__VA_ARGS__);     \
      return -1;                     \
    }                                \
  } while (0)

#define DK_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) {                                                                       \
      dk_set_error("%s:%d: kernel launch failed -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return -3;                                                                                   \
    }                                                                                              \
    (ctx)->launches++;                                                                             \
  } while (0)

// Tiled tensor map over a 16-bit tensor, 128B swizzle, zero OOB fill.
//   rank 2..4; dims[0] is the contiguous dimension; strides_bytes has rank-1 entries (dims 1..).
//   elem_strides (optional, rank entries): traversal stride per dimension; to load N elements with stride s the box
//   entry must be N * s (cuTensorMapEncodeTiled semantics).
int dk_make_tmap_16b(dk_ctx* ctx, CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr);

static inline int dk_ceil_div(int a, int b) { return (a + b - 1) / b; }
