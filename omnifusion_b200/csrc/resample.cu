// Spherical resamplers: equi2pers gather, pers2equi CSR blend, confidence merge.
// HBM-bound gather kernels: one thread per output element, tables read once per
// thread and reused across the batch, outputs written coalesced.
#include "common.cuh"

namespace ofb {

// ------------------------------------------------------------------ equi2pers
// Bilinear taps of F.grid_sample(bilinear, border, align_corners=True)
// (equi_pers/equi2pers_v3.py:111).  Written with explicit round-to-nearest
// intrinsics so no FMA contraction can change the integer taps: the result must be
// bit-identical to ATen's CPU kernel, which evaluates (g + 1) * ((size - 1) / 2),
// clamps to [0, size-1] and floors.
struct E2PTaps {
  int x0, y0;
  float wx, wy;
};

__device__ __forceinline__ E2PTaps e2p_taps(float gx, float gy, int He, int We) {
  float ix = __fmul_rn(__fadd_rn(gx, 1.f), 0.5f * (float)(We - 1));
  float iy = __fmul_rn(__fadd_rn(gy, 1.f), 0.5f * (float)(He - 1));
  ix = fminf(fmaxf(ix, 0.f), (float)(We - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(He - 1));
  float fx = floorf(ix), fy = floorf(iy);
  E2PTaps t;
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  t.wx = __fsub_rn(ix, fx);
  t.wy = __fsub_rn(iy, fy);
  return t;
}

__device__ __forceinline__ float e2p_sample(const float* __restrict__ plane, const E2PTaps& t,
                                            int He, int We) {
  const float* r0 = plane + (size_t)t.y0 * We + t.x0;
  bool xin = t.x0 + 1 < We, yin = t.y0 + 1 < He;
  float nw = __ldg(r0);
  float ne = xin ? __ldg(r0 + 1) : 0.f;
  float sw = yin ? __ldg(r0 + We) : 0.f;
  float se = (xin && yin) ? __ldg(r0 + We + 1) : 0.f;
  float ex = 1.f - t.wx, ey = 1.f - t.wy;
  return nw * (ex * ey) + ne * (t.wx * ey) + sw * (ex * t.wy) + se * (t.wx * t.wy);
}

// REF layout: out[b][c][i][j][n].  One thread per (i,j,n), n fastest -> coalesced stores.
__global__ void e2p_ref_kernel(const float* __restrict__ erp, const float2* __restrict__ grid,
                               float* __restrict__ out, int B, int C, int He, int We, int N,
                               int Ph, int Pw) {
  int total = N * Ph * Pw;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  int n = s % N;
  int ij = s / N;
  float2 g = __ldg(&grid[(size_t)n * Ph * Pw + ij]);
  E2PTaps t = e2p_taps(g.x, g.y, He, We);
  size_t plane = (size_t)He * We;
  for (int bc = 0; bc < B * C; ++bc)
    out[(size_t)bc * total + s] = e2p_sample(erp + bc * plane, t, He, We);
}

// FOLDED layout: out[b*N+n][i][j][Cpad].  One thread per (n,i,j) and chunk of E2P_EB panoramas (blockIdx.y): the
// sampling grid entry and the tap arithmetic live in registers and are reused for every panorama / channel of the
// chunk, whose 4 * C * E2P_EB gathers are all independent and in flight together.  (A thread used to walk the whole
// batch: one wave of long-running CTAs, half of the SMs idle behind the slow polar patches.)
constexpr int E2P_EB = 2;

template <int C>
__global__ void __launch_bounds__(256)
e2p_folded_kernel(const float* __restrict__ erp, const float2* __restrict__ grid,
                                  float* __restrict__ out, int B, int He, int We, int N, int Ph,
                                  int Pw) {
  int total = N * Ph * Pw;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  float2 g = __ldg(&grid[s]);
  E2PTaps t = e2p_taps(g.x, g.y, He, We);
  size_t plane = (size_t)He * We;
  const int b0 = blockIdx.y * E2P_EB;
#pragma unroll
  for (int k = 0; k < E2P_EB; ++k) {
    const int b = b0 + k;
    if (b >= B) break;
    const float* img = erp + (size_t)b * C * plane;
    if (C == 3) {
      float4 v;
      v.x = e2p_sample(img, t, He, We);
      v.y = e2p_sample(img + plane, t, He, We);
      v.z = e2p_sample(img + 2 * plane, t, He, We);
      v.w = 0.f;
      st4(out + ((size_t)b * total + s) * 4, v);
    } else {
      for (int c = 0; c < C; ++c)
        out[((size_t)b * total + s) * C + c] = e2p_sample(img + c * plane, t, He, We);
    }
  }
}

// STEM16 layout: split-half planes of (B*N, Ph, Pw+8, 4); the 4-pixel row pads are never written.
__global__ void __launch_bounds__(256)
e2p_stem16_kernel(const float* __restrict__ erp, const float2* __restrict__ grid,
                                  __half* __restrict__ out, int B, int He, int We, int N, int Ph, int Pw) {
  int total = N * Ph * Pw;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  float2 g = __ldg(&grid[s]);
  E2PTaps t = e2p_taps(g.x, g.y, He, We);
  size_t plane = (size_t)He * We;
  int j = s % Pw, ni = s / Pw;                       // ni = n*Ph + i
  const int pitch = Pw + 8;
  const size_t out_plane = (size_t)B * N * Ph * pitch * 4;
  const int b0 = blockIdx.y * E2P_EB;
  float4 v[E2P_EB];
#pragma unroll
  for (int k = 0; k < E2P_EB; ++k) {                 // all gathers of the chunk first ...
    const float* img = erp + (size_t)min(b0 + k, B - 1) * 3 * plane;
    v[k].x = e2p_sample(img, t, He, We);
    v[k].y = e2p_sample(img + plane, t, He, We);
    v[k].z = e2p_sample(img + 2 * plane, t, He, We);
    v[k].w = 0.f;
  }
#pragma unroll
  for (int k = 0; k < E2P_EB; ++k) {                 // ... then the stores
    const int b = b0 + k;
    if (b >= B) break;
    size_t o = (((size_t)b * N * Ph + ni) * pitch + 4 + j) * 4;
    act_st4<true>(out, o, out_plane, v[k]);
  }
}

__global__ void e2p_taps_kernel(const float2* __restrict__ grid, int total, int He, int We,
                                int32_t* __restrict__ x0, int32_t* __restrict__ y0) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  float2 g = grid[s];
  E2PTaps t = e2p_taps(g.x, g.y, He, We);
  x0[s] = t.x0;
  y0[s] = t.y0;
}

// ------------------------------------------------------------------ pers2equi
struct P2EStrides {
  long long sb, sc, sy, sx, sn;
};

__device__ __forceinline__ void p2e_decode(uint32_t id, int& n, int& y0, int& x0, int& dy, int& dx) {
  n = id >> 24;
  y0 = (id >> 16) & 255;
  x0 = (id >> 8) & 255;
  dy = (id >> 1) & 1;
  dx = id & 1;
}

// Both blend kernels share one structure.  A CTA owns BL_TILE consecutive ERP pixels (thread = pixel).  Their CSR
// rows are one contiguous range of the table, so the CTA copies that range - packed tap indices and pre-normalised
// weight vectors, 20 bytes per (pixel, covering patch) - into shared memory ONCE with fully coalesced loads, and
// then walks the whole batch reading the table from shared memory: per panorama only the 2x2 tap gathers touch
// global memory.  (Before: every thread walked its own row in global memory, once per group of four panoramas -
// three dependent global round trips per entry and the table re-fetched B/4 times.)  PL panoramas are gathered
// together so 4*PL (8*PL) independent loads are in flight per table entry.  Entries beyond the staged capacity
// (only possible where one tile is covered by very many patches) are read from global memory.
// Tap/weight order follows pers2equi_v3.py:174-177,194-196.
constexpr int BL_TILE = 256;
constexpr int BL_CAP = 2048;          // staged entries: 8 KB of indices + 32 KB of weights

struct BlendStage {
  uint32_t idx[BL_CAP];
  float4 w[BL_CAP];
  int row[BL_TILE + 1];
};

// returns false for threads beyond the last pixel; beg/end are relative to e0
__device__ __forceinline__ bool blend_stage(BlendStage& st, const int32_t* __restrict__ rowptr,
                                            const uint32_t* __restrict__ idx, const float4* __restrict__ w, int npix,
                                            int& e0, int& beg, int& end) {
  const int t = threadIdx.x, p0 = blockIdx.x * BL_TILE;
  st.row[t] = __ldg(&rowptr[min(p0 + t, npix)]);
  if (t == 0) st.row[BL_TILE] = __ldg(&rowptr[min(p0 + BL_TILE, npix)]);
  __syncthreads();
  e0 = st.row[0];
  const int cnt = min(st.row[BL_TILE] - e0, BL_CAP);
  for (int i = t; i < cnt; i += BL_TILE) {
    st.idx[i] = __ldg(&idx[e0 + i]);
    st.w[i] = __ldg(&w[e0 + i]);
  }
  __syncthreads();
  beg = st.row[t] - e0;
  end = st.row[t + 1] - e0;
  return p0 + t < npix;
}

template <int PL>
__global__ void __launch_bounds__(BL_TILE)
p2e_kernel(const float* __restrict__ pers, const int32_t* __restrict__ rowptr, const uint32_t* __restrict__ idx,
           const float4* __restrict__ w, float* __restrict__ out, int npix, int B, int C, P2EStrides st) {
  __shared__ BlendStage sm;
  int e0, beg, end;
  if (!blend_stage(sm, rowptr, idx, w, npix, e0, beg, end)) return;
  const int pix = blockIdx.x * BL_TILE + threadIdx.x;
  const int planes = B * C;
  for (int plane0 = 0; plane0 < planes; plane0 += PL) {
    const float* base[PL];
#pragma unroll
    for (int k = 0; k < PL; ++k) {
      const int q = min(plane0 + k, planes - 1);
      base[k] = pers + (q / C) * st.sb + (q % C) * st.sc;
    }
    float acc[PL];
#pragma unroll
    for (int k = 0; k < PL; ++k) acc[k] = 0.f;
    for (int e = beg; e < end; ++e) {
      int n, y0, x0, dy, dx;
      const bool staged = e < BL_CAP;
      p2e_decode(staged ? sm.idx[e] : __ldg(&idx[e0 + e]), n, y0, x0, dy, dx);
      const float4 ww = staged ? sm.w[e] : __ldg(&w[e0 + e]);
      const long long o00 = y0 * st.sy + x0 * st.sx + n * st.sn;
      const long long oy = dy * st.sy, ox = dx * st.sx;
#pragma unroll
      for (int k = 0; k < PL; ++k) {
        const float* p = base[k] + o00;
        const float a = __ldg(p), b = __ldg(p + oy), c = __ldg(p + ox), d = __ldg(p + oy + ox);
        acc[k] += a * ww.x + b * ww.y + c * ww.z + d * ww.w;
      }
    }
#pragma unroll
    for (int k = 0; k < PL; ++k)
      if (plane0 + k < planes) out[(size_t)(plane0 + k) * npix + pix] = acc[k];
  }
}

// Fused confidence merge (spherical_model_iterative.py:372-378): blends pred*w and w with the same table walk and
// divides.  IL: the two patch maps arrive interleaved per pixel, (pred*w, w) as one float2 (the layout the
// tensor-core heads kernel writes), so every tap is ONE 8-byte load instead of two 4-byte loads from two arrays.
template <int PL, bool IL>
__global__ void __launch_bounds__(BL_TILE)
blend_conf_kernel(const float* __restrict__ pred, const float* __restrict__ conf, const int32_t* __restrict__ rowptr,
                  const uint32_t* __restrict__ idx, const float4* __restrict__ w, float* __restrict__ out, int npix,
                  int B, int N, int Ph, int Pw) {
  __shared__ BlendStage sm;
  int e0, beg, end;
  if (!blend_stage(sm, rowptr, idx, w, npix, e0, beg, end)) return;
  const int pix = blockIdx.x * BL_TILE + threadIdx.x;
  const size_t img = (size_t)Ph * Pw;
  const float2* pc = reinterpret_cast<const float2*>(pred);
  for (int b0 = 0; b0 < B; b0 += PL) {
    float accD[PL], accW[PL];
#pragma unroll
    for (int k = 0; k < PL; ++k) accD[k] = accW[k] = 0.f;
    for (int e = beg; e < end; ++e) {
      int n, y0, x0, dy, dx;
      const bool staged = e < BL_CAP;
      p2e_decode(staged ? sm.idx[e] : __ldg(&idx[e0 + e]), n, y0, x0, dy, dx);
      const float4 ww = staged ? sm.w[e] : __ldg(&w[e0 + e]);
      const size_t o00 = (size_t)n * img + y0 * Pw + x0;
      const int oy = dy * Pw, ox = dx;
#pragma unroll
      for (int k = 0; k < PL; ++k) {
        const size_t o = (size_t)min(b0 + k, B - 1) * N * img + o00;
        if (IL) {
          const float2 a = __ldg(pc + o), b = __ldg(pc + o + oy), c = __ldg(pc + o + ox), d = __ldg(pc + o + oy + ox);
          accD[k] += a.x * ww.x + b.x * ww.y + c.x * ww.z + d.x * ww.w;
          accW[k] += a.y * ww.x + b.y * ww.y + c.y * ww.z + d.y * ww.w;
        } else {
          const float* p = pred + o;
          const float* q = conf + o;
          accD[k] += __ldg(p) * ww.x + __ldg(p + oy) * ww.y + __ldg(p + ox) * ww.z + __ldg(p + oy + ox) * ww.w;
          accW[k] += __ldg(q) * ww.x + __ldg(q + oy) * ww.y + __ldg(q + ox) * ww.z + __ldg(q + oy + ox) * ww.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PL; ++k)
      if (b0 + k < B) {
        const float W = accW[k];
        const float zero = (W <= 1e-8f) ? 1.f : 0.f;
        out[(size_t)(b0 + k) * npix + pix] = accD[k] / (W + 1e-8f * zero);
      }
  }
}

// component `comp` of an interleaved pair map -> contiguous (debug / test read-back of the engine's head outputs)
__global__ void deinterleave_kernel(const float2* __restrict__ src, size_t n, int comp, float* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 v = src[i];
  dst[i] = comp ? v.y : v.x;
}

// ------------------------------------------------------------------ backward of the two resamplers
// Training direction (supervision back-propagates through both resamplers).  Both forwards are
// linear in their image argument with input-independent taps and weights, so the backward is the transposed
// gather: a scatter-add of the incoming gradient with the same weights (atomicAdd; the summation order, and with
// it the last bits, is not deterministic - as in ATen's grid_sampler backward).
// equi2pers: grad_erp[b,c,y_t,x_t] += w_t * grad_pers[b,c,i,j,n] over the 4 bilinear taps (REF layout).
__global__ void e2p_backward_kernel(const float* __restrict__ gpers, const float2* __restrict__ grid,
                                    float* __restrict__ gerp, int B, int C, int He, int We, int N, int Ph, int Pw) {
  const int total = N * Ph * Pw;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= total) return;
  const int n = s % N, ij = s / N;
  const float2 g = __ldg(&grid[(size_t)n * Ph * Pw + ij]);
  const E2PTaps t = e2p_taps(g.x, g.y, He, We);
  const bool xin = t.x0 + 1 < We, yin = t.y0 + 1 < He;
  const float ex = 1.f - t.wx, ey = 1.f - t.wy;
  const float w00 = ex * ey, w01 = t.wx * ey, w10 = ex * t.wy, w11 = t.wx * t.wy;
  const size_t plane = (size_t)He * We;
  for (int bc = 0; bc < B * C; ++bc) {
    const float gv = __ldg(&gpers[(size_t)bc * total + s]);
    float* r0 = gerp + (size_t)bc * plane + (size_t)t.y0 * We + t.x0;
    atomicAdd(r0, gv * w00);
    if (xin) atomicAdd(r0 + 1, gv * w01);
    if (yin) atomicAdd(r0 + We, gv * w10);
    if (xin && yin) atomicAdd(r0 + We + 1, gv * w11);
  }
}

// pers2equi: grad_pers[b,c,tap,n] += w_e[tap] * grad_erp[b,c,pix] over the CSR entries of the pixel.
__global__ void __launch_bounds__(BL_TILE)
p2e_backward_kernel(const float* __restrict__ gerp, const int32_t* __restrict__ rowptr, const uint32_t* __restrict__ idx,
                    const float4* __restrict__ w, float* __restrict__ gpers, int npix, int B, int C, P2EStrides st) {
  __shared__ BlendStage sm;
  int e0, beg, end;
  if (!blend_stage(sm, rowptr, idx, w, npix, e0, beg, end)) return;
  const int pix = blockIdx.x * BL_TILE + threadIdx.x;
  const int planes = B * C;
  for (int q = 0; q < planes; ++q) {
    const float gv = __ldg(&gerp[(size_t)q * npix + pix]);
    float* base = gpers + (q / C) * st.sb + (q % C) * st.sc;
    for (int e = beg; e < end; ++e) {
      int n, y0, x0, dy, dx;
      const bool staged = e < BL_CAP;
      p2e_decode(staged ? sm.idx[e] : __ldg(&idx[e0 + e]), n, y0, x0, dy, dx);
      const float4 ww = staged ? sm.w[e] : __ldg(&w[e0 + e]);
      float* p = base + y0 * st.sy + x0 * st.sx + n * st.sn;
      const long long oy = dy * st.sy, ox = dx * st.sx;
      atomicAdd(p, gv * ww.x);
      atomicAdd(p + oy, gv * ww.y);
      atomicAdd(p + ox, gv * ww.z);
      atomicAdd(p + oy + ox, gv * ww.w);
    }
  }
}

// ------------------------------------------------------------------- abs-rel
__global__ void absrel_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                              const uint8_t* __restrict__ mask, size_t n, float scale,
                              double* __restrict__ out) {
  double s = 0.0, c = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (mask[i]) {
      float g = gt[i];
      s += (double)(fabsf(pred[i] * scale - g) / g);
      c += 1.0;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __shared__ double ss[32], sc[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { ss[wid] = s; sc[wid] = c; }
  __syncthreads();
  if (wid == 0) {
    int nw = blockDim.x >> 5;
    s = lane < nw ? ss[lane] : 0.0;
    c = lane < nw ? sc[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], c); }
  }
}

// All seven metrics of metrics.py:7-26 / test.py:151-170 in one pass:
// out = [sum |p-g|/g, sum (p-g)^2/g, sum (p-g)^2, sum (log p - log g)^2, n_log, n(d<1.25), n(d<1.25^2), n(d<1.25^3), n]
__global__ void depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                     const uint8_t* __restrict__ mask, size_t n, float scale,
                                     const float* __restrict__ scale_dev, double* __restrict__ out) {
  if (scale_dev) scale = __ldg(scale_dev);
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!mask[i]) continue;
    float g = gt[i], p = pred[i] * scale;
    float d = p - g;
    acc[0] += (double)(fabsf(d) / g);
    acc[1] += (double)((d * d) / g);
    acc[2] += (double)(d * d);
    if (p > 1e-7f && g > 1e-7f) {
      float l = logf(p) - logf(g);
      acc[3] += (double)(l * l);
      acc[4] += 1.0;
    }
    float r = fmaxf(p / g, g / p);
    acc[5] += r < 1.25f ? 1.0 : 0.0;
    acc[6] += r < 1.5625f ? 1.0 : 0.0;
    acc[7] += r < 1.953125f ? 1.0 : 0.0;
    acc[8] += 1.0;
  }
  __shared__ double sh[9][32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[k][wid] = v;
  }
  __syncthreads();
  if (wid == 0) {
    int nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      double v = lane < nw ? sh[k][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(&out[k], v);
    }
  }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_depth_metrics_partial(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                                         float scale, double* out, void* stream) {
  OFB_CHECK(pred && gt && mask && out, "depth_metrics: null pointer");
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  depth_metrics_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, mask, n, scale, nullptr, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_depth_metrics_partial_ds(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                                            const float* scale_dev, double* out, void* stream) {
  OFB_CHECK(pred && gt && mask && out && scale_dev, "depth_metrics: null pointer");
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  depth_metrics_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, mask, n, 1.f, scale_dev, out);
  OFB_LAUNCH_CHECK();
  return 0;
}


extern "C" int ofb_equi2pers_f32(const float* erp, int B, int C, int He, int We, const float* grid,
                                 int N, int Ph, int Pw, float* out, int layout, void* stream) {
  OFB_CHECK(erp && grid && out, "equi2pers: null pointer");
  OFB_CHECK(B > 0 && C > 0 && He > 1 && We > 1 && N > 0 && Ph > 0 && Pw > 0, "equi2pers: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  int total = N * Ph * Pw;
  int thr = 256, blocks = cdiv(total, thr);
  const float2* g = reinterpret_cast<const float2*>(grid);
  if (layout == OFB_LAYOUT_REF) {
    e2p_ref_kernel<<<blocks, thr, 0, s>>>(erp, g, out, B, C, He, We, N, Ph, Pw);
  } else if (layout == OFB_LAYOUT_FOLDED) {
    const dim3 grd(blocks, cdiv(B, E2P_EB));
    if (C == 3) e2p_folded_kernel<3><<<grd, thr, 0, s>>>(erp, g, out, B, He, We, N, Ph, Pw);
    else if (C == 1) e2p_folded_kernel<1><<<grd, thr, 0, s>>>(erp, g, out, B, He, We, N, Ph, Pw);
    else OFB_CHECK(false, "equi2pers: folded layout supports C in {1,3}, got %d", C);
  } else if (layout == OFB_LAYOUT_STEM16) {
    OFB_CHECK(C == 3, "equi2pers: the stem layout needs C == 3, got %d", C);
    const dim3 grd(blocks, cdiv(B, E2P_EB));
    e2p_stem16_kernel<<<grd, thr, 0, s>>>(erp, g, reinterpret_cast<__half*>(out), B, He, We, N, Ph, Pw);
  } else {
    OFB_CHECK(false, "equi2pers: unknown layout %d", layout);
  }
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_equi2pers_taps(const float* grid, int N, int Ph, int Pw, int He, int We,
                                  int32_t* x0, int32_t* y0, void* stream) {
  OFB_CHECK(grid && x0 && y0, "equi2pers_taps: null pointer");
  int total = N * Ph * Pw;
  e2p_taps_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(grid), total, He, We, x0, y0);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_pers2equi_f32(const float* pers, int B, int C, int N, int Ph, int Pw, int layout,
                                 const int32_t* rowptr, const uint32_t* idx, const float* w, int He,
                                 int We, float* out, void* stream) {
  OFB_CHECK(pers && rowptr && idx && w && out, "pers2equi: null pointer");
  OFB_CHECK(B > 0 && C > 0 && N > 0 && N <= 256 && Ph <= 256 && Pw <= 256, "pers2equi: bad shape");
  P2EStrides st;
  if (layout == OFB_LAYOUT_REF) {
    st.sn = 1; st.sx = N; st.sy = (long long)Pw * N; st.sc = (long long)Ph * Pw * N; st.sb = st.sc * C;
  } else if (layout == OFB_LAYOUT_FOLDED) {
    st.sc = 1; st.sx = C; st.sy = (long long)Pw * C; st.sn = (long long)Ph * Pw * C; st.sb = st.sn * N;
  } else {
    OFB_CHECK(false, "pers2equi: unknown layout %d", layout);
  }
  const int npix = He * We, planes = B * C;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = cdiv(npix, BL_TILE);
  if (planes >= 4) p2e_kernel<4><<<grid, BL_TILE, 0, s>>>(pers, rowptr, idx, w4, out, npix, B, C, st);
  else if (planes >= 2) p2e_kernel<2><<<grid, BL_TILE, 0, s>>>(pers, rowptr, idx, w4, out, npix, B, C, st);
  else p2e_kernel<1><<<grid, BL_TILE, 0, s>>>(pers, rowptr, idx, w4, out, npix, B, C, st);
  OFB_LAUNCH_CHECK();
  return 0;
}

namespace ofb {
int blend_conf_launch(const float* pred_w, const float* conf, bool interleaved, int B, int N, int Ph, int Pw,
                      const int32_t* rowptr, const uint32_t* idx, const float* w, int He, int We, float* out,
                      cudaStream_t s) {
  const int npix = He * We;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const int grid = cdiv(npix, BL_TILE);
  if (interleaved) {
    if (B >= 4) blend_conf_kernel<4, true><<<grid, BL_TILE, 0, s>>>(pred_w, nullptr, rowptr, idx, w4, out, npix, B, N, Ph, Pw);
    else blend_conf_kernel<1, true><<<grid, BL_TILE, 0, s>>>(pred_w, nullptr, rowptr, idx, w4, out, npix, B, N, Ph, Pw);
  } else {
    if (B >= 4) blend_conf_kernel<4, false><<<grid, BL_TILE, 0, s>>>(pred_w, conf, rowptr, idx, w4, out, npix, B, N, Ph, Pw);
    else blend_conf_kernel<1, false><<<grid, BL_TILE, 0, s>>>(pred_w, conf, rowptr, idx, w4, out, npix, B, N, Ph, Pw);
  }
  OFB_LAUNCH_CHECK();
  return 0;
}
int deinterleave_launch(const float* src_pairs, size_t n, int comp, float* dst, cudaStream_t s) {
  deinterleave_kernel<<<cdiv((long long)n, 256), 256, 0, s>>>(reinterpret_cast<const float2*>(src_pairs), n, comp, dst);
  OFB_LAUNCH_CHECK();
  return 0;
}
}  // namespace ofb

extern "C" int ofb_blend_conf_f32(const float* pred_w, const float* conf, int B, int N, int Ph, int Pw,
                                  const int32_t* rowptr, const uint32_t* idx, const float* w, int He,
                                  int We, float* out, void* stream) {
  OFB_CHECK(pred_w && conf && rowptr && idx && w && out, "blend_conf: null pointer");
  return blend_conf_launch(pred_w, conf, false, B, N, Ph, Pw, rowptr, idx, w, He, We, out, (cudaStream_t)stream);
}

extern "C" int ofb_blend_conf_pairs_f32(const float* pred_conf, int B, int N, int Ph, int Pw, const int32_t* rowptr,
                                        const uint32_t* idx, const float* w, int He, int We, float* out,
                                        void* stream) {
  OFB_CHECK(pred_conf && rowptr && idx && w && out, "blend_conf_pairs: null pointer");
  OFB_CHECK((reinterpret_cast<uintptr_t>(pred_conf) & 7) == 0, "blend_conf_pairs: the pair map must be 8-byte aligned");
  return blend_conf_launch(pred_conf, nullptr, true, B, N, Ph, Pw, rowptr, idx, w, He, We, out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ loader-side input conversion
// dataset_loader_stanford.py:52,79: rgb.astype(np.float32) / 255 and transpose(2,0,1) of the decoded uint8
// panorama (cv2 channel order kept, as the reference keeps it).  One thread per pixel; the division is the IEEE
// float32 division numpy performs, so the result is bit-identical.
__global__ void u8hwc_to_f32chw_kernel(const uint8_t* __restrict__ src, size_t pixels_per_img, int C, size_t total,
                                       float* __restrict__ dst) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / pixels_per_img, px = i - b * pixels_per_img;
  for (int c = 0; c < C; ++c)
    dst[(b * C + c) * pixels_per_img + px] = __fdiv_rn((float)src[i * C + c], 255.f);
}

extern "C" int ofb_u8hwc_to_f32chw(const uint8_t* src, int B, int H, int W, int C, float* dst, void* stream) {
  OFB_CHECK(src && dst && B > 0 && H > 0 && W > 0 && C >= 1 && C <= 4, "u8hwc_to_f32chw: bad arguments");
  const size_t ppi = (size_t)H * W, total = ppi * B;
  u8hwc_to_f32chw_kernel<<<(unsigned)cdiv((long long)total, 256), 256, 0, (cudaStream_t)stream>>>(src, ppi, C, total, dst);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_absrel_partial(const float* pred, const float* gt, const uint8_t* mask, size_t n,
                                  float scale, double* out, void* stream) {
  OFB_CHECK(pred && gt && mask && out, "absrel: null pointer");
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  absrel_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, mask, n, scale, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_equi2pers_backward_f32(const float* grad_pers, int B, int C, int He, int We, const float* grid, int N,
                                          int Ph, int Pw, float* grad_erp, void* stream) {
  OFB_CHECK(grad_pers && grid && grad_erp, "equi2pers_backward: null pointer");
  OFB_CHECK(B > 0 && C > 0 && He > 1 && We > 1 && N > 0 && Ph > 0 && Pw > 0, "equi2pers_backward: bad shape");
  const int total = N * Ph * Pw;
  e2p_backward_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_pers, reinterpret_cast<const float2*>(grid),
                                                                       grad_erp, B, C, He, We, N, Ph, Pw);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_pers2equi_backward_f32(const float* grad_erp, int B, int C, int N, int Ph, int Pw, const int32_t* rowptr,
                                          const uint32_t* idx, const float* w, int He, int We, float* grad_pers,
                                          void* stream) {
  OFB_CHECK(grad_erp && rowptr && idx && w && grad_pers, "pers2equi_backward: null pointer");
  OFB_CHECK(B > 0 && C > 0 && N > 0 && N <= 256 && Ph <= 256 && Pw <= 256, "pers2equi_backward: bad shape");
  P2EStrides st;           // reference layout (B,C,Ph,Pw,N)
  st.sn = 1; st.sx = N; st.sy = (long long)Pw * N; st.sc = (long long)Ph * Pw * N; st.sb = st.sc * C;
  const int npix = He * We;
  p2e_backward_kernel<<<cdiv(npix, BL_TILE), BL_TILE, 0, (cudaStream_t)stream>>>(
      grad_erp, rowptr, idx, reinterpret_cast<const float4*>(w), grad_pers, npix, B, C, st);
  OFB_LAUNCH_CHECK();
  return 0;
}
