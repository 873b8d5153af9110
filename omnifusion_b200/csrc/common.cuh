// Shared helpers for libofb (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/ofb.h"

namespace ofb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define OFB_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      ofb::set_error(__VA_ARGS__);           \
      return -1;                             \
    }                                        \
  } while (0)

#define OFB_CUDA(expr)                                                           \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ofb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                     __FILE__, __LINE__);                                        \
      return -2;                                                                 \
    }                                                                            \
  } while (0)

#define OFB_LAUNCH_CHECK()                                                       \
  do {                                                                           \
    ofb::count_launch();                                                         \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ofb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                        \
      return -3;                                                                 \
    }                                                                            \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace ofb
