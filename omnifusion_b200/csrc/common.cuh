// Shared helpers for libofb (sm_100a).
#pragma once
#include <cuda_fp16.h>
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

// activation element access for the two storage formats (include/ofb.h: OFB_FMT_*)
template <bool SPLIT>
__device__ __forceinline__ float4 act_ld4(const void* base, size_t idx, size_t plane) {
  if (!SPLIT) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
  const __half* h = reinterpret_cast<const __half*>(base) + idx;
  uint2 a = __ldg(reinterpret_cast<const uint2*>(h));
  uint2 b = __ldg(reinterpret_cast<const uint2*>(h + plane));
  float2 a0 = __half22float2(*reinterpret_cast<__half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<__half2*>(&a.y));
  float2 b0 = __half22float2(*reinterpret_cast<__half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<__half2*>(&b.y));
  return make_float4(a0.x + b0.x, a0.y + b0.y, a1.x + b1.x, a1.y + b1.y);
}
template <bool SPLIT>
__device__ __forceinline__ void act_st4(void* base, size_t idx, size_t plane, float4 v) {
  if (!SPLIT) { st4(reinterpret_cast<float*>(base) + idx, v); return; }
  __half* h = reinterpret_cast<__half*>(base) + idx;
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  uint2 a, b;
  a.x = *reinterpret_cast<uint32_t*>(&h0); a.y = *reinterpret_cast<uint32_t*>(&h1);
  b.x = *reinterpret_cast<uint32_t*>(&l0); b.y = *reinterpret_cast<uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(h) = a;
  *reinterpret_cast<uint2*>(h + plane) = b;
}


// scalar variants
template <bool SPLIT>
__device__ __forceinline__ float act_ld1(const void* base, size_t idx, size_t plane) {
  if (!SPLIT) return __ldg(reinterpret_cast<const float*>(base) + idx);
  const __half* h = reinterpret_cast<const __half*>(base) + idx;
  return __half2float(__ldg(h)) + __half2float(__ldg(h + plane));
}
template <bool SPLIT>
__device__ __forceinline__ void act_st1(void* base, size_t idx, size_t plane, float v) {
  if (!SPLIT) { reinterpret_cast<float*>(base)[idx] = v; return; }
  __half* h = reinterpret_cast<__half*>(base) + idx;
  __half hi = __float2half_rn(v);
  h[0] = hi;
  h[plane] = __float2half_rn(v - __half2float(hi));
}

}  // namespace ofb
