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

// tcgen05 launch variants (csrc/conv_tc.cu), owned by an engine handle: "cta2" / "pdl" / "store128" / "fill_div" /
// "direct32" / "khr_bw" / "khr_row64" of ofb_set_option, plus sm_share (two-lane forwards) and the timing-experiment
// switches (dbg).
struct TcOptions {
  bool pdl = true;        // programmatic dependent launch
  bool store128 = true;   // bulk-tensor-store epilogue also for the 128-wide tiles
  bool cta2 = true;       // cta_group::2 CTA pairs for the 128-wide split-half tiles
  bool direct32 = false;  // BN = 32 split-half tiles store straight from registers
  bool khr_row64 = false; // 64-byte K rows for the Cout = 64 kh-reuse layers
  bool wmc = false;       // kh-reuse kernels: clusters of two CTAs share each stage's weight loads by TMA multicast
                          // (measured: the weight share of the L2 -> SM traffic - 37 % for 128 -> 32 @64x64 - is halved,
                          // the launch times do not move: these layers are bound by MMA issue, not by L2)
  bool nstack = true;     // heads kernel: input-row-stationary MMAs with the three kh taps stacked along N
  bool nstack_ups = false;// the same for the fused-upsample conv (measured slower: experiments only)
  int fill_div = 2;       // shrink the N tile while fewer than num_sms / fill_div tiles exist
  int khr_bw = 16;        // tile width of the kh-reuse kernels (16 or 32)
  int sm_share = 1;       // persistent grids use num_sms / sm_share CTAs
  int dbg = 0;            // TcParams::dbg (timing experiments; results are wrong)
};
const TcOptions& tc_opts();
struct TcOptScope {
  const TcOptions* prev;
  explicit TcOptScope(const TcOptions* o);
  ~TcOptScope();
  TcOptScope(const TcOptScope&) = delete;
  TcOptScope& operator=(const TcOptScope&) = delete;
};

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


// 8-channel variants (two float4 / two 16-byte half loads)
struct float8 { float4 a, b; };
template <bool SPLIT>
__device__ __forceinline__ float8 act_ld8(const void* base, size_t idx, size_t plane) {
  float8 r;
  if (!SPLIT) {
    const float* f = reinterpret_cast<const float*>(base) + idx;
    r.a = __ldg(reinterpret_cast<const float4*>(f));
    r.b = __ldg(reinterpret_cast<const float4*>(f + 4));
    return r;
  }
  const __half* h = reinterpret_cast<const __half*>(base) + idx;
  uint4 x = __ldg(reinterpret_cast<const uint4*>(h));
  uint4 y = __ldg(reinterpret_cast<const uint4*>(h + plane));
  const __half2* xh = reinterpret_cast<const __half2*>(&x);
  const __half2* yh = reinterpret_cast<const __half2*>(&y);
  float2 p0 = __half22float2(xh[0]), p1 = __half22float2(xh[1]), p2 = __half22float2(xh[2]), p3 = __half22float2(xh[3]);
  float2 q0 = __half22float2(yh[0]), q1 = __half22float2(yh[1]), q2 = __half22float2(yh[2]), q3 = __half22float2(yh[3]);
  r.a = make_float4(p0.x + q0.x, p0.y + q0.y, p1.x + q1.x, p1.y + q1.y);
  r.b = make_float4(p2.x + q2.x, p2.y + q2.y, p3.x + q3.x, p3.y + q3.y);
  return r;
}
template <bool SPLIT>
__device__ __forceinline__ void act_st8(void* base, size_t idx, size_t plane, const float8& v) {
  if (!SPLIT) {
    float* f = reinterpret_cast<float*>(base) + idx;
    st4(f, v.a);
    st4(f + 4, v.b);
    return;
  }
  __half* h = reinterpret_cast<__half*>(base) + idx;
  const float in[8] = {v.a.x, v.a.y, v.a.z, v.a.w, v.b.x, v.b.y, v.b.z, v.b.w};
  uint4 hi4, lo4;
  __half2* hh = reinterpret_cast<__half2*>(&hi4);
  __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    __half2 x = __floats2half2_rn(in[2 * t], in[2 * t + 1]);
    float2 xf = __half22float2(x);
    hh[t] = x;
    ll[t] = __floats2half2_rn(in[2 * t] - xf.x, in[2 * t + 1] - xf.y);
  }
  *reinterpret_cast<uint4*>(h) = hi4;
  *reinterpret_cast<uint4*>(h + plane) = lo4;
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
