// Bandwidth-bound network kernels over folded NHWC: stem conv, max-pool, 2x bilinear
// upsample, point embedding, token packing, LayerNorm, attention core, depth/confidence heads.
#include "common.cuh"
#include "ptx.cuh"

namespace ofb {

// ----------------------------------------------------------------------- stem
// Conv 7x7 s2 p3, 3(+1 pad)->64, BN, ReLU (spherical_model_iterative.py:322).
// CTA: 16x16 output pixels x 64 couts; 128 threads, each 4 pixels (rows r, r+4, r+8, r+12) x 32
// couts = 128 accumulators, so every weight read from shared memory feeds 4 FMAs per channel.
// Input halo tile (37x37 float4) and the whole 50 KB filter live in shared memory.
constexpr int STEM_TH = 16, STEM_TW = 16;
constexpr int STEM_IH = STEM_TH * 2 + 5, STEM_IW = STEM_TW * 2 + 5;
constexpr int STEM_SMEM = (49 * 4 * 64 + STEM_IH * STEM_IW * 4) * 4;

template <bool OUT_SPLIT>
__global__ void __launch_bounds__(128)
stem_kernel(const float* __restrict__ in, int n, int h, int w, const float* __restrict__ wgt,
            const float* __restrict__ scale, const float* __restrict__ shift, void* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;                               // [tap][c][64]
  float4* si = reinterpret_cast<float4*>(smem + 49 * 4 * 64);  // [IH][IW]
  const int oh_n = h / 2, ow_n = w / 2;
  const int tiles_w = ow_n / STEM_TW, tiles_h = oh_n / STEM_TH;
  int t = blockIdx.x;
  int img = t / (tiles_w * tiles_h);
  int r = t - img * tiles_w * tiles_h;
  int th = r / tiles_w, tw = r - th * tiles_w;
  int oh0 = th * STEM_TH, ow0 = tw * STEM_TW;
  int ih0 = oh0 * 2 - 3, iw0 = ow0 * 2 - 3;
  const int tid = threadIdx.x;
  // weights (7,7,4,64) arrive with one 50 KB bulk copy issued by thread 0 (TMA engine), overlapping
  // the cooperative halo-tile fill below
  __shared__ __align__(8) uint64_t wbar;
  const uint32_t bar = smem_u32(&wbar);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, 49 * 4 * 64 * 4);
    bulk_load(smem_u32(sw), wgt, 49 * 4 * 64 * 4, bar);
  }
  {  // halo tile: all of a thread's loads are issued before the first shared-memory store
    constexpr int PER = (STEM_IH * STEM_IW + 127) / 128;
    float4 v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      int i = tid + u * 128;
      int y = i / STEM_IW, x = i - y * STEM_IW;
      int ih = ih0 + y, iw = iw0 + x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < STEM_IH * STEM_IW && ih >= 0 && ih < h && iw >= 0 && iw < w)
        v[u] = __ldg(reinterpret_cast<const float4*>(in + ((size_t)(img * h + ih) * w + iw) * 4));
    }
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (tid + u * 128 < STEM_IH * STEM_IW) si[tid + u * 128] = v[u];
  }
  __syncthreads();
  mbar_wait(bar, 0);
  int half = tid >> 6;                 // cout half: 32 couts
  int p = tid & 63;                    // pixel quad id
  int pr = p / STEM_TW, pc = p - pr * STEM_TW;   // pr in [0,4)
  float acc[4][32];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[q][j] = 0.f;
  for (int kh = 0; kh < 7; ++kh) {
    for (int kw = 0; kw < 7; ++kw) {
      float4 x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = si[((pr + 4 * q) * 2 + kh) * STEM_IW + pc * 2 + kw];
      const float* wp = sw + (kh * 7 + kw) * 4 * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 w0 = *reinterpret_cast<const float4*>(wp + j);
        float4 w1 = *reinterpret_cast<const float4*>(wp + 64 + j);
        float4 w2 = *reinterpret_cast<const float4*>(wp + 128 + j);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q][j] += x[q].x * w0.x + x[q].y * w1.x + x[q].z * w2.x;
          acc[q][j + 1] += x[q].x * w0.y + x[q].y * w1.y + x[q].z * w2.y;
          acc[q][j + 2] += x[q].x * w0.z + x[q].y * w1.z + x[q].z * w2.z;
          acc[q][j + 3] += x[q].x * w0.w + x[q].y * w1.w + x[q].z * w2.w;
        }
      }
    }
  }
  const size_t plane = (size_t)n * oh_n * ow_n * 64;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const size_t o = ((size_t)(img * oh_n + oh0 + pr + 4 * q) * ow_n + ow0 + pc) * 64 + half * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 sc = __ldg(reinterpret_cast<const float4*>(scale + half * 32 + j));
      float4 b = __ldg(reinterpret_cast<const float4*>(shift + half * 32 + j));
      float4 v;
      v.x = fmaxf(acc[q][j] * sc.x + b.x, 0.f); v.y = fmaxf(acc[q][j + 1] * sc.y + b.y, 0.f);
      v.z = fmaxf(acc[q][j + 2] * sc.z + b.z, 0.f); v.w = fmaxf(acc[q][j + 3] * sc.w + b.w, 0.f);
      act_st4<OUT_SPLIT>(out, o + j, plane, v);
    }
  }
}

// -------------------------------------------------------------------- maxpool
// F.max_pool3d((3,3,1), s(2,2,1), p(1,1,0)).  Thread = (image, band of output rows, output column, 8-channel
// chunk); it walks its band top to bottom and carries the horizontal 3-max of input row 2*oh+1 over to the next
// output row (where it is row 2*oh'-1), so every input row is fetched once instead of 1.5 times and an output costs
// 6 instead of 9 taps; all loads are 16 bytes per plane, a warp touches whole 128-byte pixel vectors.
template <bool SPLIT>
__device__ __forceinline__ float8 pool_hmax(const void* __restrict__ in, size_t row_off, int ow, int w, int c8, int c,
                                            size_t plane) {
  float8 m;
  m.a = m.b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dx = -1; dx <= 1; ++dx) {
    const int iw = ow * 2 + dx;
    if (iw < 0 || iw >= w) continue;
    const float8 v = act_ld8<SPLIT>(in, (row_off + iw) * c8 * 8 + (size_t)c * 8, plane);
    m.a.x = fmaxf(m.a.x, v.a.x); m.a.y = fmaxf(m.a.y, v.a.y); m.a.z = fmaxf(m.a.z, v.a.z); m.a.w = fmaxf(m.a.w, v.a.w);
    m.b.x = fmaxf(m.b.x, v.b.x); m.b.y = fmaxf(m.b.y, v.b.y); m.b.z = fmaxf(m.b.z, v.b.z); m.b.w = fmaxf(m.b.w, v.b.w);
  }
  return m;
}
__device__ __forceinline__ float8 max8(const float8& x, const float8& y) {
  float8 m;
  m.a = make_float4(fmaxf(x.a.x, y.a.x), fmaxf(x.a.y, y.a.y), fmaxf(x.a.z, y.a.z), fmaxf(x.a.w, y.a.w));
  m.b = make_float4(fmaxf(x.b.x, y.b.x), fmaxf(x.b.y, y.b.y), fmaxf(x.b.z, y.b.z), fmaxf(x.b.w, y.b.w));
  return m;
}

template <bool SPLIT>
__global__ void __launch_bounds__(256)
maxpool_kernel(const void* __restrict__ in, int n, int h, int w, int c8, int bands, void* __restrict__ out) {
  const int oh_n = h / 2, ow_n = w / 2;
  // 32-bit index arithmetic (the host checks the range)
  const uint32_t total = (uint32_t)n * bands * ow_n * c8;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (uint32_t)c8);
  uint32_t r = i / (uint32_t)c8;
  const int ow = (int)(r % (uint32_t)ow_n); r /= (uint32_t)ow_n;
  const int band = (int)(r % (uint32_t)bands);
  const int img = (int)(r / (uint32_t)bands);
  const int rows = (oh_n + bands - 1) / bands;
  const int oh0 = band * rows, oh1 = min(oh0 + rows, oh_n);
  const size_t in_plane = (size_t)n * h * w * c8 * 8, out_plane = (size_t)n * oh_n * ow_n * c8 * 8;
  float8 carry;                                   // horizontal max of input row 2*oh - 1
  carry.a = carry.b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  if (oh0 > 0) carry = pool_hmax<SPLIT>(in, ((size_t)img * h + 2 * oh0 - 1) * w, ow, w, c8, c, in_plane);
  for (int oh = oh0; oh < oh1; ++oh) {
    const float8 mid = pool_hmax<SPLIT>(in, ((size_t)img * h + 2 * oh) * w, ow, w, c8, c, in_plane);
    const float8 low = pool_hmax<SPLIT>(in, ((size_t)img * h + 2 * oh + 1) * w, ow, w, c8, c, in_plane);   // 2*oh+1 < h: h is even
    const float8 m = max8(max8(carry, mid), low);
    act_st8<SPLIT>(out, ((((size_t)img * oh_n + oh) * ow_n + ow) * c8 + c) * 8, out_plane, m);
    carry = low;
  }
}

// ------------------------------------------------------------------- upsample
// F.interpolate(scale 2, bilinear, align_corners=False): src = (dst+0.5)/2-0.5 clamped at 0,
// i1 = min(i0+1, size-1), lambda in {0, .25, .75}.  Same expression tree as ATen's
// upsample_bilinear2d: h0*(w0*a + w1*b) + h1*(w0*c + w1*d).
// One thread per (source pixel, 8-channel chunk): it loads the 3x3 source neighbourhood once
// (clamped at the borders) and produces the 2x2 output block, i.e. 2.25 loads per output instead of 4.
// Per output the expression tree is ATen's: h0*(w0*a + w1*b) + h1*(w0*c + w1*d).
template <bool SPLIT>
__global__ void __launch_bounds__(256, 3)
upsample2x_kernel(const void* __restrict__ in, const float* __restrict__ img_bias,
                                  int n, int h, int w, int c8, void* __restrict__ out) {
  const uint32_t total = (uint32_t)n * h * w * c8;       // 32-bit index arithmetic, see maxpool_kernel
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (uint32_t)c8);
  uint32_t r = i / (uint32_t)c8;
  const int x = (int)(r % (uint32_t)w); r /= (uint32_t)w;
  const int y = (int)(r % (uint32_t)h);
  const int img = (int)(r / (uint32_t)h);
  const size_t in_plane = (size_t)n * h * w * c8 * 8, out_plane = in_plane * 4;
  const int ys[3] = {max(y - 1, 0), y, min(y + 1, h - 1)};
  const int xs[3] = {max(x - 1, 0), x, min(x + 1, w - 1)};
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (img_bias) {
    const float4* bp = reinterpret_cast<const float4*>(img_bias) + ((size_t)img * c8 + c) * 2;
    b0 = __ldg(bp); b1 = __ldg(bp + 1);
  }
  auto load_row = [&](int a, float8* row) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      row[b] = act_ld8<SPLIT>(in, (((size_t)img * h + ys[a]) * w + xs[b]) * c8 * 8 + (size_t)c * 8, in_plane);
      if (img_bias) {
        row[b].a.x += b0.x; row[b].a.y += b0.y; row[b].a.z += b0.z; row[b].a.w += b0.w;
        row[b].b.x += b1.x; row[b].b.y += b1.y; row[b].b.z += b1.z; row[b].b.w += b1.w;
      }
    }
  };
  // Output row 2y+dy blends source rows (dy, dy+1) of the clamped 3x3 block; same for columns.
  // dy = 0: rows (y-1, y) weighted (0.25, 0.75); at the top border both are row 0 and the weights
  // become (0, 1) so the value is exact, like ATen's lambda = 0.  dy = 1: rows (y, y+1) with (0.75, 0.25).
  // Two source rows are live at a time (the third replaces the first): 48 instead of 72 data registers.
  float8 ra[3], rb[3];
  load_row(0, ra);
  load_row(1, rb);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    if (dy == 1) load_row(2, ra);                  // rows (1, 2): rb is the upper one now
    const float8* up = dy == 0 ? ra : rb;
    const float8* dn = dy == 0 ? rb : ra;
    const float h1 = dy == 0 ? (y == 0 ? 1.f : 0.75f) : 0.25f, h0 = 1.f - h1;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float w1 = dx == 0 ? (x == 0 ? 1.f : 0.75f) : 0.25f, w0 = 1.f - w1;
      const float8 &A = up[dx], &B = up[dx + 1], &Cc = dn[dx], &D = dn[dx + 1];
#define OFB_UP(f) (h0 * (w0 * A.f + w1 * B.f) + h1 * (w0 * Cc.f + w1 * D.f))
      float8 o;
      o.a = make_float4(OFB_UP(a.x), OFB_UP(a.y), OFB_UP(a.z), OFB_UP(a.w));
      o.b = make_float4(OFB_UP(b.x), OFB_UP(b.y), OFB_UP(b.z), OFB_UP(b.w));
#undef OFB_UP
      size_t o_idx = ((((size_t)img * 2 * h + 2 * y + dy) * 2 * w + 2 * x + dx) * c8 + c) * 8;
      act_st8<SPLIT>(out, o_idx, out_plane, o);
    }
  }
}

// ---------------------------------------------------------------- point embed
// Thread = 4 pixels x 8 of the 64 output channels.  The work per pixel is a 16 x 64 matrix-vector product whose
// operand (the second-layer weights, shared memory) must be re-read for every pixel: with one pixel per thread the
// kernel is bound by the shared-memory pipe (measured: 85 % LSU wavefront utilisation at a quarter of the HBM rate).
// Four pixels per thread reuse every 16-byte weight load for 4 x 4 FMAs; the 16 hidden units of a pixel (<= 5 inputs
// each, uniform weights = broadcast loads) are recomputed by the 8 threads that share the pixel, which is cheaper
// than exchanging them.  A warp covers 16 consecutive pixels: every base load / result store instruction moves four
// whole 256-byte pixel vectors.  Accumulation order as in the reference's conv: c, then k ascending.
constexpr int PE_PIX = 4;

template <bool SPLIT>
__global__ void __launch_bounds__(256)
point_embed_kernel(const float* __restrict__ pts, int N, int cin, int p,
                                   const float* __restrict__ depth, int imgs,
                                   const float* __restrict__ w1, const float* __restrict__ s1,
                                   const float* __restrict__ t1, const float* __restrict__ w2,
                                   const float* __restrict__ s2, const float* __restrict__ t2,
                                   const void* __restrict__ base, void* __restrict__ out) {
  __shared__ __align__(16) float w2s[16][64];
  __shared__ __align__(16) float w1s[16][8];       // [k][c0..c4, scale, shift, -]
  for (int i = threadIdx.x; i < 1024; i += 256) w2s[i & 15][i >> 4] = __ldg(&w2[i]);   // w2 is [co][k]
  if (threadIdx.x < 128) {
    const int k = threadIdx.x >> 3, c = threadIdx.x & 7;
    w1s[k][c] = c < cin ? __ldg(&w1[k * cin + c]) : (c == 5 ? __ldg(&s1[k]) : (c == 6 ? __ldg(&t1[k]) : 0.f));
  }
  __syncthreads();
  const uint32_t npix = (uint32_t)imgs * p * p;             // 32-bit index arithmetic (the host checks the range)
  const size_t plane = (size_t)npix * 64;
  const int g = threadIdx.x & 7, sub = (threadIdx.x & 31) >> 3;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(&s2[g * 8 + j]); sh[j] = __ldg(&t2[g * 8 + j]); }
  const uint32_t pp = (uint32_t)(p * p);
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5, warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (uint32_t blk = warp0; blk * (4 * PE_PIX) < npix; blk += warps) {       // 16 pixels per warp and step
    float v[PE_PIX][5];
    uint32_t pix[PE_PIX];
#pragma unroll
    for (int q = 0; q < PE_PIX; ++q) {
      pix[q] = blk * (4 * PE_PIX) + q * 4 + sub;
      const uint32_t pc = min(pix[q], npix - 1);
      const int xy = (int)(pc % pp);
      const int n = (int)((pc / pp) % (uint32_t)N);
      const float dsc = depth ? __ldg(&depth[pc]) : 1.f;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        v[q][c] = c < cin ? __ldg(&pts[((size_t)n * cin + c) * pp + xy]) : 0.f;
        if (depth && c < 3) v[q][c] *= dsc;
      }
    }
    float o[PE_PIX][8];
#pragma unroll
    for (int q = 0; q < PE_PIX; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[q][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const float4 wl = *reinterpret_cast<const float4*>(&w1s[k][0]);
      const float4 wh = *reinterpret_cast<const float4*>(&w1s[k][4]);      // c4, scale, shift
      const float4 wa = *reinterpret_cast<const float4*>(&w2s[k][g * 8]);
      const float4 wb = *reinterpret_cast<const float4*>(&w2s[k][g * 8 + 4]);
#pragma unroll
      for (int q = 0; q < PE_PIX; ++q) {
        float a = 0.f;
        a += v[q][0] * wl.x;
        if (cin > 1) a += v[q][1] * wl.y;
        if (cin > 2) a += v[q][2] * wl.z;
        if (cin > 3) a += v[q][3] * wl.w;
        if (cin > 4) a += v[q][4] * wh.x;
        const float hk = fmaxf(a * wh.y + wh.z, 0.f);
        o[q][0] += hk * wa.x; o[q][1] += hk * wa.y; o[q][2] += hk * wa.z; o[q][3] += hk * wa.w;
        o[q][4] += hk * wb.x; o[q][5] += hk * wb.y; o[q][6] += hk * wb.z; o[q][7] += hk * wb.w;
      }
    }
#pragma unroll
    for (int q = 0; q < PE_PIX; ++q) {
      if (pix[q] >= npix) continue;
      float8 r;
      r.a = make_float4(fmaxf(o[q][0] * sc[0] + sh[0], 0.f), fmaxf(o[q][1] * sc[1] + sh[1], 0.f),
                        fmaxf(o[q][2] * sc[2] + sh[2], 0.f), fmaxf(o[q][3] * sc[3] + sh[3], 0.f));
      r.b = make_float4(fmaxf(o[q][4] * sc[4] + sh[4], 0.f), fmaxf(o[q][5] * sc[5] + sh[5], 0.f),
                        fmaxf(o[q][6] * sc[6] + sh[6], 0.f), fmaxf(o[q][7] * sc[7] + sh[7], 0.f));
      const size_t off = (size_t)pix[q] * 64 + g * 8;
      if (base) {
        const float8 bb = act_ld8<SPLIT>(base, off, plane);
        r.a.x += bb.a.x; r.a.y += bb.a.y; r.a.z += bb.a.z; r.a.w += bb.a.w;
        r.b.x += bb.b.x; r.b.y += bb.b.y; r.b.z += bb.b.z; r.b.w += bb.b.w;
      }
      act_st8<SPLIT>(out, off, plane, r);
    }
  }
}

// ----------------------------------------------------------------- token pack
// down (imgs, S, S, Cd) with S*S*Cd == 512 -> tokens (imgs, 512), token index = c*S*S + i*S + j (the reference's
// reshape of the (B, Cd, S, S, N) map, spherical_model_iterative.py:330-331), + pos_emb.  S = 4, Cd = 32 for 128x128
// patches; S = 8, Cd = 8 for the 256x256 variant (network_test.py:271).
template <bool SPLIT, bool OUT_SPLIT>
__global__ void token_pack_kernel(const void* __restrict__ down, const float* __restrict__ pos,
                                  int imgs, int N, int ss, int cstride, void* __restrict__ tokens) {
  // cstride: channels per position of `down` in memory (>= 512 / ss: narrower reductions are padded to 32 channels)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= imgs * 512) return;
  int img = i >> 9, t = i & 511;
  int c = t / ss, ij = t - c * ss;
  const size_t in_plane = (size_t)imgs * ss * cstride, plane = (size_t)imgs * 512;
  float v = act_ld1<SPLIT>(down, ((size_t)img * ss + ij) * cstride + c, in_plane) + __ldg(&pos[(img % N) * 512 + t]);
  act_st1<OUT_SPLIT>(tokens, i, plane, v);
}

// ------------------------------------------------------------------ layernorm
// One warp per row; mean, then variance of the centred values (two passes in registers).
template <int DIM, bool IN_SPLIT, bool OUT_SPLIT>
__global__ void layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int rows, float eps,
                                 void* __restrict__ y) {
  constexpr int PER = DIM / 32;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t plane = (size_t)rows * DIM;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    float4 t = act_ld4<IN_SPLIT>(x, (size_t)row * DIM + (i / 4) * 128 + lane * 4, plane);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    s += t.x + t.y + t.z + t.w;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / DIM;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  float rstd = rsqrtf(q / DIM + eps);
  // rsqrtf is approximate (2 ulp); one Newton step makes it correctly-rounded-ish
  float var = q / DIM + eps;
  rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    int col = (i / 4) * 128 + lane * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i] - mean) * rstd * g.x + b.x;
    o.y = (v[i + 1] - mean) * rstd * g.y + b.y;
    o.z = (v[i + 2] - mean) * rstd * g.z + b.z;
    o.w = (v[i + 3] - mean) * rstd * g.w + b.w;
    act_st4<OUT_SPLIT>(y, (size_t)row * DIM + col, plane, o);
  }
}

// ------------------------------------------------------ split-K finish + LayerNorm
// The two 512-wide linears of a Transformer_Block (attn.proj, mlp.fc2; model/blocks.py:84-88) run split-K
// on the tcgen05 engine (few token rows: the K range is what can be spread over the SMs) and leave raw
// float32 partial sums.  One warp per row: x_new = residual + bias + wscale * sum_s partial[s] (fixed order),
// written as the new residual stream, and LayerNorm(x_new) - the input of the next linear - in the same pass,
// so the finish replaces the LayerNorm launch instead of adding one.
template <bool LN_SPLIT>
__global__ void splitk_ln_kernel(const float* __restrict__ partial, int ksplit, float wscale,
                                 const float* __restrict__ bias, const void* __restrict__ residual, int rows,
                                 void* __restrict__ x_out, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, void* __restrict__ ln_out) {
  constexpr int DIM = 512, PER = DIM / 32;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t plane = (size_t)rows * DIM;
  float v[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    const int col = (i / 4) * 128 + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < ksplit; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)k * rows + row) * DIM + col));
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    const float4 r = act_ld4<true>(residual, (size_t)row * DIM + col, plane);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + col));
    // same expression as the conv epilogue: acc * scale + shift, then + residual
    float4 o;
    o.x = (acc.x * wscale + b.x) + r.x; o.y = (acc.y * wscale + b.y) + r.y;
    o.z = (acc.z * wscale + b.z) + r.z; o.w = (acc.w * wscale + b.w) + r.w;
    act_st4<true>(x_out, (size_t)row * DIM + col, plane, o);
    // LayerNorm reads what the next kernel would read back: the split-half rounded value
    const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2half2_rn(o.z - f1.x, o.w - f1.y);
    const float2 g0 = __half22float2(l0), g1 = __half22float2(l1);
    v[i] = f0.x + g0.x; v[i + 1] = f0.y + g0.y; v[i + 2] = f1.x + g1.x; v[i + 3] = f1.y + g1.y;
    s += v[i] + v[i + 1] + v[i + 2] + v[i + 3];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / DIM;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float var = q / DIM + eps;
  float rstd = rsqrtf(var);
  rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    const int col = (i / 4) * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i] - mean) * rstd * g.x + b.x;
    o.y = (v[i + 1] - mean) * rstd * g.y + b.y;
    o.z = (v[i + 2] - mean) * rstd * g.z + b.z;
    o.w = (v[i + 3] - mean) * rstd * g.w + b.w;
    act_st4<LN_SPLIT>(ln_out, (size_t)row * DIM + col, plane, o);
  }
}

// ------------------------------------------------------------------ attention
// One CTA per (panorama, head, group of ATT_RG query rows): N <= 64 tokens, head_dim 128.  K, V and the
// group's Q rows are staged in shared memory (rows padded to 132 floats: 16-byte aligned, bank-skewed),
// scores + softmax + PV entirely on chip.  Splitting the query rows over CTAs triples the number of CTAs
// of this latency-bound kernel (32 -> 96 at 8 panoramas).
constexpr int ATT_MAXN = 64, ATT_D = 128, ATT_LD = ATT_D + 4, ATT_RG = 6;

template <bool SPLIT>
__global__ void __launch_bounds__(128)
attention_kernel(const void* __restrict__ q, int q_ld, const void* __restrict__ kv, int kv_ld, int kv_col0,
                 int N, int heads, int groups, float scale, void* __restrict__ out, int rows) {
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                       // [ATT_RG][ATT_LD]
  float* sk = sq + ATT_RG * ATT_LD;     // [N][ATT_LD]
  float* sv = sk + N * ATT_LD;          // [N][ATT_LD]
  float* sp = sv + N * ATT_LD;          // [ATT_RG][N+1]
  const int g = blockIdx.x % groups, bh = blockIdx.x / groups;
  const int b = bh / heads, hd = bh % heads;
  const int r_begin = g * ATT_RG, nr = min(ATT_RG, N - r_begin);
  const int dim = heads * ATT_D;
  const int tid = threadIdx.x;
  // fill: 8 channels per load, four independent loads in flight per thread before the smem stores
  {
    const int qchunks = nr * 16, chunks = qchunks + 2 * N * 16;     // (row, 8-channel chunk) of q | k | v
    const size_t qplane = (size_t)rows * q_ld, kvplane = (size_t)rows * kv_ld;
    for (int i0 = tid; i0 < chunks; i0 += 128 * 4) {
      float8 v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 128;
        dst[u] = -1;
        if (i < chunks) {
          if (i < qchunks) {
            const int r = i >> 4, d = (i & 15) * 8;
            v[u] = act_ld8<SPLIT>(q, ((size_t)b * N + r_begin + r) * q_ld + hd * ATT_D + d, qplane);
            dst[u] = r * ATT_LD + d;
          } else {
            const int j = i - qchunks;
            const int which = j / (N * 16), rem = j - which * (N * 16);
            const int r = rem >> 4, d = (rem & 15) * 8;
            v[u] = act_ld8<SPLIT>(kv, ((size_t)b * N + r) * kv_ld + kv_col0 + which * dim + hd * ATT_D + d, kvplane);
            dst[u] = (ATT_RG + which * N + r) * ATT_LD + d;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (dst[u] < 0) continue;
        *reinterpret_cast<float4*>(sm + dst[u]) = v[u].a;
        *reinterpret_cast<float4*>(sm + dst[u] + 4) = v[u].b;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nr * N; i += 128) {
    const int r = i / N, c = i % N;
    const float4* qr = reinterpret_cast<const float4*>(sq + r * ATT_LD);
    const float4* kr = reinterpret_cast<const float4*>(sk + c * ATT_LD);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int d = 0; d < ATT_D / 4; ++d) {
      const float4 x = qr[d], y = kr[d];
      a0 += x.x * y.x; a1 += x.y * y.y; a2 += x.z * y.z; a3 += x.w * y.w;
    }
    sp[r * (N + 1) + c] = ((a0 + a1) + (a2 + a3)) * scale;
  }
  __syncthreads();
  // softmax: one warp per row
  const int lane = tid & 31, wid = tid >> 5;
  for (int r = wid; r < nr; r += 4) {
    float m = -INFINITY;
    for (int c = lane; c < N; c += 32) m = fmaxf(m, sp[r * (N + 1) + c]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < N; c += 32) {
      float e = expf(sp[r * (N + 1) + c] - m);
      sp[r * (N + 1) + c] = e;
      s += e;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float inv = 1.f / s;
    for (int c = lane; c < N; c += 32) sp[r * (N + 1) + c] *= inv;
  }
  __syncthreads();
  // out[r][d] = sum_c P[r][c] V[c][d]; thread = d
  float a[ATT_RG];
#pragma unroll
  for (int u = 0; u < ATT_RG; ++u) a[u] = 0.f;
  for (int c = 0; c < N; ++c) {
    const float vv = sv[c * ATT_LD + tid];
#pragma unroll
    for (int u = 0; u < ATT_RG; ++u) a[u] += sp[min(u, nr - 1) * (N + 1) + c] * vv;
  }
#pragma unroll
  for (int u = 0; u < ATT_RG; ++u)
    if (u < nr) act_st1<SPLIT>(out, ((size_t)b * N + r_begin + u) * dim + hd * ATT_D + tid, (size_t)rows * dim, a[u]);
}

// ---------------------------------------------------------------------- heads
// pred / weight_pred 3x3 32->1 convs sharing one read of de_conv4_0
// (spherical_model_iterative.py:371-374).  CTA = 16x16 pixels; the 18x18x32 halo tile is kept
// pixel-major with a 36-float pitch: 16-byte loads/stores are bank-conflict free both for the
// cooperative fill and for the per-pixel reads (pitch 36 -> lane l starts at bank 4l).
constexpr int HD_T = 16, HD_I = HD_T + 2, HD_PITCH = 36;
constexpr int HD_SMEM = HD_I * HD_I * HD_PITCH * 4;
// Both 3x3x32 filters live in constant memory: every lane of a warp needs the same weight at
// the same time, so the FFMAs take it as a constant-bank operand and no shared-memory bandwidth
// is spent on weights (the kernel was shared-memory bound with the filters in smem).
__constant__ float c_heads_w[2][288];   // [pred | conf][tap][32]

template <bool SPLIT>
__global__ void __launch_bounds__(256)
heads_kernel(const void* __restrict__ x, int imgs, int h, int w, const float* __restrict__ wp,
             float bp, const float* __restrict__ wc, float bc, int confidence,
             float* __restrict__ pred_out, float* __restrict__ conf_out) {
  extern __shared__ __align__(16) float hsm[];
  float* tile = hsm;
  int tiles_w = w / HD_T, tiles_h = h / HD_T;
  int t = blockIdx.x;
  int img = t / (tiles_w * tiles_h);
  int r = t - img * tiles_w * tiles_h;
  int y0 = (r / tiles_w) * HD_T - 1, x0 = (r % tiles_w) * HD_T - 1;
  int tid = threadIdx.x;
  const size_t xplane = (size_t)imgs * h * w * 32;
  {  // halo tile, 4 channels per load; all loads in flight before the first shared-memory store
    constexpr int TOTAL = HD_I * HD_I * 8, PER = (TOTAL + 255) / 256;
    float4 v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      int i = tid + u * 256;
      int cq = i & 7, pix = i >> 3;
      int yy = pix / HD_I, xx = pix - yy * HD_I;
      int ih = y0 + yy, iw = x0 + xx;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < TOTAL && ih >= 0 && ih < h && iw >= 0 && iw < w)
        v[u] = act_ld4<SPLIT>(x, ((size_t)(img * h + ih) * w + iw) * 32 + cq * 4, xplane);
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      int i = tid + u * 256;
      if (i < TOTAL) *reinterpret_cast<float4*>(tile + (i >> 3) * HD_PITCH + (i & 7) * 4) = v[u];
    }
  }
  __syncthreads();
  int py = tid / HD_T, px = tid % HD_T;
  float ap = 0.f, ac = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float* tp = tile + ((py + ky) * HD_I + px + kx) * HD_PITCH;
      const int wo = (ky * 3 + kx) * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        float4 v = *reinterpret_cast<const float4*>(tp + c);
        ap += v.x * c_heads_w[0][wo + c] + v.y * c_heads_w[0][wo + c + 1] + v.z * c_heads_w[0][wo + c + 2] +
              v.w * c_heads_w[0][wo + c + 3];
        ac += v.x * c_heads_w[1][wo + c] + v.y * c_heads_w[1][wo + c + 1] + v.z * c_heads_w[1][wo + c + 2] +
              v.w * c_heads_w[1][wo + c + 3];
      }
    }
  size_t o = (size_t)(img * h + y0 + 1 + py) * w + x0 + 1 + px;
  float pred = fmaxf(ap + bp, 0.f);
  if (confidence) {
    float cf = 1.f / (1.f + expf(-(ac + bc)));
    pred_out[o] = pred * cf;
    conf_out[o] = cf;
  } else {
    pred_out[o] = pred;
  }
}

// ------------------------------------------------------------------ stem row pads
// The tensor-core stem reads patches in the STEM16 layout (split-half planes of (imgs, P, P + 8, 4)): every row carries
// 4 zero pixels on each side, which equi2pers never writes.  They are re-zeroed by EVERY forward inside the
// stream-ordered (hence graph-captured) region, because the arena is re-planned per batch size and another forward
// may have put live data there.  The right pad of row r and the left pad of row r + 1 are one contiguous 64-byte
// span (rows of both planes are back to back): one thread per row boundary.
__global__ void zero_stem_pads_kernel(uint4* __restrict__ base, uint32_t rows, uint32_t pitch16) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;        // boundary after row r - 1 (r == 0: head, r == rows: tail)
  if (r > rows) return;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint4* p = base + (size_t)r * pitch16;
  if (r > 0) { p[-2] = z; p[-1] = z; }
  if (r < rows) { p[0] = z; p[1] = z; }
}

int zero_stem_pads(void* patches, int imgs, int P, cudaStream_t s) {
  const uint32_t rows = (uint32_t)2 * imgs * P, pitch16 = (uint32_t)(P + 8) * 4 * 2 / 16;
  zero_stem_pads_kernel<<<cdiv((long long)rows + 1, 256), 256, 0, s>>>(reinterpret_cast<uint4*>(patches), rows, pitch16);
  OFB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------- range check
// max |x| and the number of non-finite elements of an activation (option "check_range"): the split-half format stores
// value = fp16 hi + fp16 lo, so |x| > 65504 overflows to inf and |x| < ~6e-5 loses the lo plane to fp16 subnormals;
// the reference is fp32 and has neither limit.  out[0] = bit pattern of max |x| (atomicMax on the non-negative float),
// out[1] = non-finite count.
__global__ void range_kernel(const void* __restrict__ p, size_t n, int fmt, unsigned int* __restrict__ out) {
  float m = 0.f;
  unsigned int bad = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v;
    if (fmt == OFB_FMT_SPLIT16) {
      const __half* h = reinterpret_cast<const __half*>(p);
      v = __half2float(h[i]) + __half2float(h[n + i]);
    } else {
      v = reinterpret_cast<const float*>(p)[i];
    }
    if (!isfinite(v)) ++bad; else m = fmaxf(m, fabsf(v));
  }
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], __float_as_uint(m));
    if (bad) atomicAdd(&out[1], bad);
  }
}

int range_launch(const void* p, size_t n, int fmt, unsigned int* out2, cudaStream_t s) {
  int blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  range_kernel<<<blocks, 256, 0, s>>>(p, n, fmt, out2);
  OFB_LAUNCH_CHECK();
  return 0;
}

}  // namespace ofb

using namespace ofb;

#define OFB_FMT_OK(f) ((f) == OFB_FMT_F32 || (f) == OFB_FMT_SPLIT16)

extern "C" int ofb_stem_f32(const float* in, int n, int h, int w, const float* wgt, const float* scale,
                            const float* shift, void* out, int out_fmt, void* stream) {
  OFB_CHECK(in && wgt && scale && shift && out && OFB_FMT_OK(out_fmt), "stem: bad arguments");
  OFB_CHECK(h % (2 * STEM_TH) == 0 && w % (2 * STEM_TW) == 0, "stem: h,w must be multiples of 32 (got %d,%d)", h, w);
  static bool attr = false;
  if (!attr) {
    OFB_CUDA(cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM));
    OFB_CUDA(cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM));
    attr = true;
  }
  int blocks = n * (h / 2 / STEM_TH) * (w / 2 / STEM_TW);
  if (out_fmt) stem_kernel<true><<<blocks, 128, STEM_SMEM, (cudaStream_t)stream>>>(in, n, h, w, wgt, scale, shift, out);
  else stem_kernel<false><<<blocks, 128, STEM_SMEM, (cudaStream_t)stream>>>(in, n, h, w, wgt, scale, shift, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_maxpool3x3s2_f32(const void* in, int n, int h, int w, int c, void* out, int fmt, void* stream) {
  OFB_CHECK(in && out && c % 8 == 0 && h % 2 == 0 && w % 2 == 0 && OFB_FMT_OK(fmt), "maxpool: bad arguments (c %% 8, even h and w)");
  // bands of output rows per image: enough threads to fill the GPU (>= ~8 warps per SM), at least 4 rows per band
  const int oh_n = h / 2;
  int bands = 1;
  while (bands * 2 <= oh_n / 4 && (size_t)n * bands * (w / 2) * (c / 8) < (size_t)148 * 2048) bands *= 2;
  size_t total = (size_t)n * bands * (w / 2) * (c / 8);
  OFB_CHECK((size_t)n * h * w * (c / 8) < (1ull << 31), "maxpool: input exceeds the 32-bit index range (process fewer images per call)");
  if (fmt) maxpool_kernel<true><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, n, h, w, c / 8, bands, out);
  else maxpool_kernel<false><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, n, h, w, c / 8, bands, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_upsample2x_f32(const void* in, const float* img_bias, int n, int h, int w, int c,
                                  void* out, int fmt, void* stream) {
  OFB_CHECK(in && out && c % 8 == 0 && OFB_FMT_OK(fmt), "upsample2x: bad arguments");
  size_t total = (size_t)n * h * w * (c / 8);
  OFB_CHECK(total < (1ull << 31), "upsample2x: %zu input vectors exceed the 32-bit index range (process fewer images per call)", total);
  if (fmt) upsample2x_kernel<true><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, img_bias, n, h, w, c / 8, out);
  else upsample2x_kernel<false><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(in, img_bias, n, h, w, c / 8, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_point_embed_f32(const float* pts, int N, int cin, int p, const float* depth, int imgs,
                                   const float* w1, const float* s1, const float* t1, const float* w2,
                                   const float* s2, const float* t2, const void* base, void* out, int fmt,
                                   void* stream) {
  OFB_CHECK(pts && w1 && s1 && t1 && w2 && s2 && t2 && out && OFB_FMT_OK(fmt), "point_embed: bad arguments");
  OFB_CHECK(cin >= 1 && cin <= 5, "point_embed: cin must be <= 5 (got %d)", cin);
  OFB_CHECK((size_t)imgs * p * p < (1ull << 31), "point_embed: too many pixels for the 32-bit index range");
  size_t total = (size_t)imgs * p * p * 8 / PE_PIX;         // threads: 8 per group of PE_PIX pixels
  // persistent CTAs (two per SM at this register count): the 4.5 KB weight stage is paid once per CTA
  const int blocks = (int)(cdiv(total, 256) < 148 * 2 ? (cdiv(total, 256) > 0 ? cdiv(total, 256) : 1) : 148 * 2);
  if (fmt) point_embed_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(pts, N, cin, p, depth, imgs, w1, s1, t1, w2, s2, t2, base, out);
  else point_embed_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(pts, N, cin, p, depth, imgs, w1, s1, t1, w2, s2, t2, base, out);
  OFB_LAUNCH_CHECK();
  return 0;
}

namespace ofb {
int token_pack(const void* down, const float* pos_emb, int imgs, int N, int spatial, int cstride, void* tokens, int fmt, cudaStream_t s, int out_fmt = -1);
int token_pack(const void* down, const float* pos_emb, int imgs, int N, int spatial, int cstride, void* tokens, int fmt, cudaStream_t s, int out_fmt) {
  OFB_CHECK(down && pos_emb && tokens && OFB_FMT_OK(fmt), "token_pack: bad arguments");
  if (out_fmt < 0) out_fmt = fmt;
  const int ss = spatial * spatial;
  OFB_CHECK(ss > 0 && 512 % ss == 0 && cstride >= 512 / ss, "token_pack: %dx%d positions x %d channels do not hold the 512-wide token", spatial, spatial, cstride);
  int blocks = cdiv((long long)imgs * 512, 256);
  if (fmt && out_fmt) token_pack_kernel<true, true><<<blocks, 256, 0, s>>>(down, pos_emb, imgs, N, ss, cstride, tokens);
  else if (fmt) token_pack_kernel<true, false><<<blocks, 256, 0, s>>>(down, pos_emb, imgs, N, ss, cstride, tokens);
  else token_pack_kernel<false, false><<<blocks, 256, 0, s>>>(down, pos_emb, imgs, N, ss, cstride, tokens);
  OFB_LAUNCH_CHECK();
  return 0;
}
}  // namespace ofb

extern "C" int ofb_token_pack_f32(const void* down, const float* pos_emb, int imgs, int N, void* tokens, int fmt,
                                  void* stream) {
  return token_pack(down, pos_emb, imgs, N, 4, 32, tokens, fmt, (cudaStream_t)stream);
}

extern "C" int ofb_layernorm_f32(const void* x, const float* gamma, const float* beta, int rows, int dim,
                                 float eps, void* y, int in_fmt, int out_fmt, void* stream) {
  OFB_CHECK(x && gamma && beta && y && OFB_FMT_OK(in_fmt) && OFB_FMT_OK(out_fmt), "layernorm: bad arguments");
  OFB_CHECK(dim == 512, "layernorm: dim must be 512 (got %d)", dim);
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = cdiv(rows, 4);
  if (!in_fmt && !out_fmt) layernorm_kernel<512, false, false><<<blocks, 128, 0, s>>>(x, gamma, beta, rows, eps, y);
  else if (in_fmt && out_fmt) layernorm_kernel<512, true, true><<<blocks, 128, 0, s>>>(x, gamma, beta, rows, eps, y);
  else if (in_fmt) layernorm_kernel<512, true, false><<<blocks, 128, 0, s>>>(x, gamma, beta, rows, eps, y);
  else layernorm_kernel<512, false, true><<<blocks, 128, 0, s>>>(x, gamma, beta, rows, eps, y);
  OFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------ split-K finish of a conv layer
// layer4's 512 -> 512 convs at 4x4 pixels per patch have K = 4608 and, at a few panoramas per step, only 18 M tiles:
// they run 2-way split-K on the tcgen05 engine (halves the operand traffic per CTA, which is what bounds them) and
// this kernel finishes them: sum of the partial sums in slice order, BN scale / shift (same expression as the conv
// epilogue), residual, ReLU, split-half store.  8 channels per thread.
__global__ void splitk_conv_finish_kernel(const float* __restrict__ partial, int ksplit, size_t m_total, int cout,
                                          const float* __restrict__ scale, const float* __restrict__ shift, float wscale,
                                          const __half* __restrict__ residual, size_t plane, int act,
                                          __half* __restrict__ out) {
  const size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 8;
  if (i >= m_total * cout) return;
  const int c = (int)(i % cout);
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  for (int s = 0; s < ksplit; ++s) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(partial + s * m_total * cout + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(partial + s * m_total * cout + i + 4));
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = f[j] * ((scale ? __ldg(scale + c + j) : 1.f) * wscale) + (shift ? __ldg(shift + c + j) : 0.f);
  if (residual) {
    const uint4 rh = __ldg(reinterpret_cast<const uint4*>(residual + i));
    const uint4 rl = __ldg(reinterpret_cast<const uint4*>(residual + plane + i));
    const __half2* ah = reinterpret_cast<const __half2*>(&rh);
    const __half2* bh = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = __half22float2(ah[t]), y = __half22float2(bh[t]);
      f[2 * t] += x.x + y.x;
      f[2 * t + 1] += x.y + y.y;
    }
  }
  if (act == OFB_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  uint4 hi4, lo4;
  __half2* hh = reinterpret_cast<__half2*>(&hi4);
  __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const __half2 h = __floats2half2_rn(f[2 * t], f[2 * t + 1]);
    const float2 hf = __half22float2(h);
    hh[t] = h;
    ll[t] = __floats2half2_rn(f[2 * t] - hf.x, f[2 * t + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(out + i) = hi4;
  *reinterpret_cast<uint4*>(out + plane + i) = lo4;
}

extern "C" int ofb_splitk_finish_conv_f16(const float* partial, int ksplit, long long pixels, int cout, const float* scale,
                                          const float* shift, float wscale, const void* residual_planes, int act,
                                          void* out_planes, void* stream) {
  OFB_CHECK(partial && out_planes && ksplit >= 1 && pixels > 0 && cout > 0 && cout % 8 == 0, "splitk_finish_conv: bad arguments");
  OFB_CHECK(act == OFB_ACT_NONE || act == OFB_ACT_RELU, "splitk_finish_conv: activation %d is not supported", act);
  const size_t n = (size_t)pixels * cout;
  const int blocks = cdiv((long long)(n / 8), 256);
  splitk_conv_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(partial, ksplit, (size_t)pixels, cout, scale, shift, wscale,
      reinterpret_cast<const __half*>(residual_planes), n, act, reinterpret_cast<__half*>(out_planes));
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_splitk_finish_ln_f32(const float* partial, int ksplit, float wscale, const float* bias,
                                        const void* residual, int rows, int dim, void* x_out, const float* gamma,
                                        const float* beta, float eps, void* ln_out, int ln_fmt, void* stream) {
  OFB_CHECK(partial && residual && x_out && gamma && beta && ln_out && OFB_FMT_OK(ln_fmt), "splitk_finish_ln: bad arguments");
  OFB_CHECK(dim == 512 && ksplit >= 1 && rows > 0, "splitk_finish_ln: dim must be 512 (got %d), ksplit >= 1", dim);
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = cdiv(rows, 4);
  if (ln_fmt) splitk_ln_kernel<true><<<blocks, 128, 0, s>>>(partial, ksplit, wscale, bias, residual, rows, x_out, gamma, beta, eps, ln_out);
  else splitk_ln_kernel<false><<<blocks, 128, 0, s>>>(partial, ksplit, wscale, bias, residual, rows, x_out, gamma, beta, eps, ln_out);
  OFB_LAUNCH_CHECK();
  return 0;
}

static int attention_launch(const void* q, int q_ld, const void* kv, int kv_ld, int kv_col0, int B, int N, int heads,
                            int head_dim, void* out, int fmt, void* stream) {
  OFB_CHECK(q && kv && out && OFB_FMT_OK(fmt), "attention: bad arguments");
  OFB_CHECK(head_dim == ATT_D && N <= ATT_MAXN && N > 0, "attention: head_dim must be 128 and N <= 64 (got %d, %d)", head_dim, N);
  const int groups = (N + ATT_RG - 1) / ATT_RG;
  int smem = ((ATT_RG + 2 * N) * ATT_LD + ATT_RG * (N + 1)) * 4;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    OFB_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    OFB_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  float sc = 1.f / sqrtf((float)head_dim);
  if (fmt) attention_kernel<true><<<B * heads * groups, 128, smem, (cudaStream_t)stream>>>(q, q_ld, kv, kv_ld, kv_col0, N, heads, groups, sc, out, B * N);
  else attention_kernel<false><<<B * heads * groups, 128, smem, (cudaStream_t)stream>>>(q, q_ld, kv, kv_ld, kv_col0, N, heads, groups, sc, out, B * N);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_attention_f32(const void* q, const void* kv, int B, int N, int heads, int head_dim,
                                 void* out, int fmt, void* stream) {
  return attention_launch(q, heads * head_dim, kv, 2 * heads * head_dim, 0, B, N, heads, head_dim, out, fmt, stream);
}

extern "C" int ofb_attention_qkv_f32(const void* qkv, int B, int N, int heads, int head_dim, void* out, int fmt,
                                     void* stream) {
  int dim = heads * head_dim;
  return attention_launch(qkv, 3 * dim, qkv, 3 * dim, dim, B, N, heads, head_dim, out, fmt, stream);
}

extern "C" int ofb_heads_f32(const void* x, int imgs, int h, int w, const float* w_pred, float b_pred,
                             const float* w_conf, float b_conf, int confidence, float* pred_out,
                             float* conf_out, int in_fmt, void* stream) {
  OFB_CHECK(x && w_pred && pred_out && (!confidence || (w_conf && conf_out)) && OFB_FMT_OK(in_fmt), "heads: bad arguments");
  OFB_CHECK(h % HD_T == 0 && w % HD_T == 0, "heads: h,w must be multiples of 16");
  int blocks = imgs * (h / HD_T) * (w / HD_T);
  // filters -> constant memory (stream-ordered device-to-device copies, 2.3 KB)
  OFB_CUDA(cudaMemcpyToSymbolAsync(c_heads_w, w_pred, 288 * sizeof(float), 0, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (confidence)
    OFB_CUDA(cudaMemcpyToSymbolAsync(c_heads_w, w_conf, 288 * sizeof(float), 288 * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  static bool attr = false;
  if (!attr) {
    OFB_CUDA(cudaFuncSetAttribute(heads_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM));
    OFB_CUDA(cudaFuncSetAttribute(heads_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_SMEM));
    attr = true;
  }
  if (in_fmt) heads_kernel<true><<<blocks, 256, HD_SMEM, (cudaStream_t)stream>>>(x, imgs, h, w, w_pred, b_pred, w_conf, b_conf, confidence, pred_out, conf_out);
  else heads_kernel<false><<<blocks, 256, HD_SMEM, (cudaStream_t)stream>>>(x, imgs, h, w, w_pred, b_pred, w_conf, b_conf, confidence, pred_out, conf_out);
  OFB_LAUNCH_CHECK();
  return 0;
}
