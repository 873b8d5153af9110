// Evaluation-side kernels around the forward: masked median by radix selection for the
// median scaling of test.py:161-162, the point cloud of test.py:205-218 / util.py:159-174, and the loader's
// INTER_AREA down-scaling (dataset_loader_stanford.py:92-96).  All HBM-bound streaming kernels.
#include "common.cuh"

namespace ofb {

// ------------------------------------------------------------------ masked median (radix select)
// torch.median returns the LOWER median: element (n-1)/2 of the sorted masked values.  Four passes over the
// data, most significant byte first: a 256-bin histogram of the current byte among the elements whose higher
// bytes equal the prefix found so far, then one tiny kernel picks the bin that contains the target rank.
// Keys are the usual order-preserving map of IEEE floats to unsigned integers, so any finite input works.
// state layout (uint32): [0,1024) histograms of the 4 passes, [1024] prefix, [1025] rank, [1026] valid count.
__device__ __forceinline__ uint32_t float_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(256)
median_hist_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, size_t n, int pass,
                   uint32_t* __restrict__ state) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t prefix = pass ? state[1024] : 0u;
  const int shift = 24 - 8 * pass;
  // 4 elements per thread and step: one 16-byte load of x, one 4-byte load of the mask
  const size_t n4 = n >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t m4 = __ldg(reinterpret_cast<const uint32_t*>(mask) + i);
    if (!m4) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!((m4 >> (8 * k)) & 0xFFu)) continue;
      const uint32_t key = float_key(vv[k]);
      if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&h[(key >> shift) & 255u], 1u);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail
    const size_t i = (n4 << 2) + threadIdx.x;
    if (mask[i]) {
      const uint32_t key = float_key(x[i]);
      if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&h[(key >> shift) & 255u], 1u);
    }
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&state[pass * 256 + threadIdx.x], h[threadIdx.x]);
}

__global__ void median_pick_kernel(int pass, uint32_t* __restrict__ state, float* __restrict__ out) {
  // one warp: inclusive scan of the 256 bins, 8 per lane
  const int lane = threadIdx.x;
  const uint32_t* h = state + pass * 256;
  uint32_t c[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k] = h[lane * 8 + k]; s += c[k]; }
  uint32_t incl = s;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t rank;
  if (pass == 0) {
    rank = total ? (total - 1) >> 1 : 0;
    if (lane == 0) state[1026] = total;
  } else {
    rank = state[1025];
  }
  uint32_t before = incl - s;
  int bin = -1;
  uint32_t rank_in = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (bin < 0 && rank < before + c[k]) { bin = lane * 8 + k; rank_in = rank - before; }
    before += c[k];
  }
  const uint32_t found = __ballot_sync(0xffffffffu, bin >= 0);
  if (!found) {                                  // empty selection: median of nothing
    if (lane == 0 && pass == 3) *out = __uint_as_float(0x7FC00000u);
    return;
  }
  const int src = __ffs(found) - 1;
  bin = __shfl_sync(0xffffffffu, bin, src);
  rank_in = __shfl_sync(0xffffffffu, rank_in, src);
  if (lane == 0) {
    const uint32_t prefix = ((pass ? state[1024] : 0u) << 8) | (uint32_t)bin;
    state[1024] = prefix;
    state[1025] = rank_in;
    if (pass == 3) *out = state[1026] ? key_float(prefix) : __uint_as_float(0x7FC00000u);
  }
}

__global__ void median_ratio_kernel(float* __restrict__ out3) { out3[0] = __fdiv_rn(out3[1], out3[2]); }

static int masked_median(const float* x, const uint8_t* mask, size_t n, uint32_t* state, float* out, cudaStream_t s) {
  OFB_CUDA(cudaMemsetAsync(state, 0, 1027 * sizeof(uint32_t), s));
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  for (int pass = 0; pass < 4; ++pass) {
    median_hist_kernel<<<blocks, 256, 0, s>>>(x, mask, n, pass, state);
    OFB_LAUNCH_CHECK();
    median_pick_kernel<<<1, 32, 0, s>>>(pass, state, out);
    OFB_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ point cloud
// test.py:205-218: pts[b, y, x, :] = rays[y, x, :] * depth[b, 0, y, x]; rays (He*We, 3) is the input-independent
// unit-ray table (util.py:159-174 coords2uv / uv2xyz, built on the host).  One thread per pixel: a 12-byte ray
// and a 4-byte depth in, 12 bytes out.
__global__ void depth_to_points_kernel(const float* __restrict__ depth, const float* __restrict__ rays, int B,
                                       uint32_t npix, float max_depth, float* __restrict__ pts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float rx = __ldg(&rays[3 * (size_t)i]), ry = __ldg(&rays[3 * (size_t)i + 1]), rz = __ldg(&rays[3 * (size_t)i + 2]);
  for (int b = 0; b < B; ++b) {
    float d = __ldg(&depth[(size_t)b * npix + i]);
    if (max_depth > 0.f && d > max_depth) d = 0.f;          // test.py:208 zeroes predictions above 8 m
    float* o = pts + ((size_t)b * npix + i) * 3;
    o[0] = __fmul_rn(rx, d); o[1] = __fmul_rn(ry, d); o[2] = __fmul_rn(rz, d);
  }
}

// ------------------------------------------------------------------ INTER_AREA resize (integer factors)
// cv2.resize(img, (W/f, H/f), interpolation=cv2.INTER_AREA) of an 8-bit HWC image with an integer scale factor f
// (dataset_loader_stanford.py:92-96 halves / quarters the 2048x4096 Stanford panoramas): every output pixel is
// the mean of its f x f source block.  OpenCV's integer-scale fast path (ResizeAreaFastVec) sums the block in
// integers and for uint8 rounds (sum * (1/f^2)) to nearest-even via saturate_cast<uchar>(float) = cvRound.
// Output: uint8 HWC (feed ofb_u8hwc_to_f32chw) - bit-identical to cv2 for f in {2, 4} (pinned by a fixture).
__global__ void area_resize_u8_kernel(const uint8_t* __restrict__ src, int H, int W, int C, int f,
                                      uint8_t* __restrict__ dst, size_t total) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int oW = W / f, oH = H / f;
  const int c = (int)(i % C);
  size_t r = i / C;
  const int ox = (int)(r % oW); r /= oW;
  const int oy = (int)(r % oH);
  const size_t b = r / oH;
  const uint8_t* p = src + ((b * H + (size_t)oy * f) * W + (size_t)ox * f) * C + c;
  uint32_t sum = 0;
  for (int dy = 0; dy < f; ++dy)
    for (int dx = 0; dx < f; ++dx) sum += p[((size_t)dy * W + dx) * C];
  if (f == 2) {
    dst[i] = (uint8_t)((sum + 2) >> 2);          // OpenCV's 2x2 uint8 fast path: (a+b+c+d+2) >> 2
  } else {
    const float v = (float)sum * (1.f / (float)(f * f));
    dst[i] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f));
  }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_masked_median_f32(const float* x, const uint8_t* mask, size_t n, void* state, float* out,
                                     void* stream) {
  OFB_CHECK(x && mask && state && out && n > 0, "masked_median: bad arguments");
  OFB_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3) == 0,
            "masked_median: x must be 16-byte and mask 4-byte aligned");
  return masked_median(x, mask, n, reinterpret_cast<uint32_t*>(state), out, (cudaStream_t)stream);
}

extern "C" int ofb_median_scale_f32(const float* pred, const float* gt, const uint8_t* mask, size_t n, void* state,
                                    float* out3, void* stream) {
  OFB_CHECK(pred && gt && mask && state && out3 && n > 0, "median_scale: bad arguments");
  OFB_CHECK(((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(mask) & 3) == 0, "median_scale: pred/gt must be 16-byte and mask 4-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* st = reinterpret_cast<uint32_t*>(state);
  if (masked_median(gt, mask, n, st, out3 + 1, s)) return -1;
  if (masked_median(pred, mask, n, st, out3 + 2, s)) return -1;
  median_ratio_kernel<<<1, 1, 0, s>>>(out3);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_depth_to_points_f32(const float* depth, const float* rays, int B, int He, int We, float max_depth,
                                       float* pts, void* stream) {
  OFB_CHECK(depth && rays && pts && B > 0 && He > 0 && We > 0, "depth_to_points: bad arguments");
  const size_t npix = (size_t)He * We;
  OFB_CHECK(npix < (1ull << 31), "depth_to_points: panorama too large");
  depth_to_points_kernel<<<cdiv((long long)npix, 256), 256, 0, (cudaStream_t)stream>>>(depth, rays, B, (uint32_t)npix, max_depth, pts);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_area_resize_u8(const uint8_t* src, int B, int H, int W, int C, int factor, uint8_t* dst,
                                  void* stream) {
  OFB_CHECK(src && dst && B > 0 && H > 0 && W > 0 && C >= 1 && C <= 4, "area_resize: bad arguments");
  OFB_CHECK(factor >= 1 && factor <= 16 && H % factor == 0 && W % factor == 0,
            "area_resize: integer scale factors that divide the image only (got %d for %dx%d)", factor, H, W);
  const size_t total = (size_t)B * (H / factor) * (W / factor) * C;
  area_resize_u8_kernel<<<(unsigned)cdiv((long long)total, 256), 256, 0, (cudaStream_t)stream>>>(src, H, W, C, factor, dst, total);
  OFB_LAUNCH_CHECK();
  return 0;
}
