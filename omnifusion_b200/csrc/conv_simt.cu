// fp32 implicit-GEMM convolution on CUDA cores (FFMA), folded NHWC.
// Used for the layers that do not map onto the tcgen05 engine (strided convs, tiny
// channel counts, nn.Linear at small row counts) and as the device-side cross-check
// for the tensor-core engine.
//
// GEMM view: M = n*oh*ow output pixels, N = cout, K = k*k*(c0+c1).  CTA tile BMxBN,
// K step 16 channels of one filter tap; A (activations) and B (OHWI weights) tiles are
// staged k-major in shared memory, double-buffered through registers; each thread owns
// a TMxTN register tile.
#include "common.cuh"

namespace ofb {

struct ConvArgs {
  const void* in0; const void* in1; int c0, c1;
  int n, h, w, oh, ow;
  const float* wgt; int k, stride, pad, cout;
  const float* scale; const float* shift; const void* residual;
  int act;
  void* out;
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

template <int BM, int BN, int TM, int TN, bool IN_SPLIT, bool OUT_SPLIT>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_simt_kernel(ConvArgs a) {
  constexpr int BK = 16;
  constexpr int THREADS = (BM / TM) * (BN / TN);
  constexpr int A_LD = (BM * 4) / THREADS;       // float4 loads per thread for the A tile
  constexpr int B_LD = (BN * 4 + THREADS - 1) / THREADS;
  static_assert((BM * 4) % THREADS == 0, "A tile must divide evenly");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "register tile must be float4-able");

  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int cin = a.c0 + a.c1;
  const int M = a.n * a.oh * a.ow;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int cchunks = cin / BK;
  const int ksteps = a.k * a.k * cchunks;

  // fixed A rows of this thread
  int a_img[A_LD], a_ih0[A_LD], a_iw0[A_LD];
  bool a_ok[A_LD];
#pragma unroll
  for (int i = 0; i < A_LD; ++i) {
    int f = tid + i * THREADS;
    int m = m0 + (f >> 2);
    a_ok[i] = m < M;
    int mm = a_ok[i] ? m : 0;
    int img = mm / (a.oh * a.ow);
    int r = mm - img * (a.oh * a.ow);
    int oh = r / a.ow, ow = r - oh * a.ow;
    a_img[i] = img;
    a_ih0[i] = oh * a.stride - a.pad;
    a_iw0[i] = ow * a.stride - a.pad;
  }

  float4 ra[A_LD], rb[B_LD];

  auto load_tiles = [&](int step) {
    int tap = step / cchunks;
    int cq = step - tap * cchunks;
    int kh = tap / a.k, kw = tap - kh * a.k;
    int c = cq * BK;
    const void* src = a.in0;
    int cs = a.c0, coff = c;
    if (c >= a.c0) { src = a.in1; cs = a.c1; coff = c - a.c0; }
    const size_t in_plane = (size_t)a.n * a.h * a.w * cs;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + i * THREADS;
      int q = f & 3;
      int ih = a_ih0[i] + kh, iw = a_iw0[i] + kw;
      bool ok = a_ok[i] && ih >= 0 && ih < a.h && iw >= 0 && iw < a.w;
      ra[i] = ok ? act_ld4<IN_SPLIT>(src, ((size_t)(a_img[i] * a.h + ih) * a.w + iw) * cs + coff + q * 4, in_plane)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int f = tid + i * THREADS;
      if (f < BN * 4) {
        int row = f >> 2, q = f & 3;
        rb[i] = __ldg(reinterpret_cast<const float4*>(
            a.wgt + ((size_t)(n0 + row) * a.k * a.k + tap) * cin + c + q * 4));
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + i * THREADS;
      int row = f >> 2, q = (f & 3) * 4;
      As[buf][q + 0][row] = ra[i].x;
      As[buf][q + 1][row] = ra[i].y;
      As[buf][q + 2][row] = ra[i].z;
      As[buf][q + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int f = tid + i * THREADS;
      if (f < BN * 4) {
        int row = f >> 2, q = (f & 3) * 4;
        Bs[buf][q + 0][row] = rb[i].x;
        Bs[buf][q + 1][row] = rb[i].y;
        Bs[buf][q + 2][row] = rb[i].z;
        Bs[buf][q + 3][row] = rb[i].w;
      }
    }
  };

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int step = 0; step < ksteps; ++step) {
    int buf = step & 1;
    if (step + 1 < ksteps) load_tiles(step + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        av[i] = v.x; av[i + 1] = v.y; av[i + 2] = v.z; av[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
        bv[j] = v.x; bv[j + 1] = v.y; bv[j + 2] = v.z; bv[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (step + 1 < ksteps) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: y = act(acc*scale + shift + residual)
  const size_t out_plane = (size_t)M * a.cout;
  float sc[TN], sh[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int co = n0 + tx * TN + j;
    sc[j] = a.scale ? __ldg(&a.scale[co]) : 1.f;
    sh[j] = a.shift ? __ldg(&a.shift[co]) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
    size_t o = (size_t)m * a.cout + n0 + tx * TN;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      float4 v;
      v.x = acc[i][j] * sc[j] + sh[j];
      v.y = acc[i][j + 1] * sc[j + 1] + sh[j + 1];
      v.z = acc[i][j + 2] * sc[j + 2] + sh[j + 2];
      v.w = acc[i][j + 3] * sc[j + 3] + sh[j + 3];
      if (a.residual) {
        float4 r = act_ld4<OUT_SPLIT>(a.residual, o + j, out_plane);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (a.act == OFB_ACT_RELU) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      } else if (a.act == OFB_ACT_GELU) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
      }
      act_st4<OUT_SPLIT>(a.out, o + j, out_plane, v);
    }
  }
}

template <int BM, int BN, int TM, int TN>
static int launch(const ConvArgs& a, int in_fmt, int out_fmt, cudaStream_t s) {
  constexpr int THREADS = (BM / TM) * (BN / TN);
  int M = a.n * a.oh * a.ow;
  dim3 grid(cdiv(M, BM), a.cout / BN);
  if (in_fmt == OFB_FMT_F32 && out_fmt == OFB_FMT_F32) conv_simt_kernel<BM, BN, TM, TN, false, false><<<grid, THREADS, 0, s>>>(a);
  else if (in_fmt == OFB_FMT_SPLIT16 && out_fmt == OFB_FMT_SPLIT16) conv_simt_kernel<BM, BN, TM, TN, true, true><<<grid, THREADS, 0, s>>>(a);
  else if (in_fmt == OFB_FMT_SPLIT16) conv_simt_kernel<BM, BN, TM, TN, true, false><<<grid, THREADS, 0, s>>>(a);
  else conv_simt_kernel<BM, BN, TM, TN, false, true><<<grid, THREADS, 0, s>>>(a);
  OFB_LAUNCH_CHECK();
  return 0;
}

int conv_simt(const ofb_conv_desc* d, cudaStream_t s) {
  ConvArgs a;
  a.in0 = d->in0; a.in1 = d->in1; a.c0 = d->c0; a.c1 = d->in1 ? d->c1 : 0;
  a.n = d->n; a.h = d->h; a.w = d->w;
  a.k = d->k; a.stride = d->stride; a.pad = d->pad; a.cout = d->cout;
  a.oh = (d->h + 2 * d->pad - d->k) / d->stride + 1;
  a.ow = (d->w + 2 * d->pad - d->k) / d->stride + 1;
  a.wgt = d->wgt; a.scale = d->scale; a.shift = d->shift; a.residual = d->residual;
  a.act = d->act; a.out = d->out;
  OFB_CHECK(a.c0 % 16 == 0 && a.c1 % 16 == 0, "conv_simt: channel counts must be multiples of 16 (got %d,%d)", a.c0, a.c1);
  OFB_CHECK(a.cout % 32 == 0, "conv_simt: cout must be a multiple of 32 (got %d)", a.cout);
  long long M = (long long)a.n * a.oh * a.ow;
  OFB_CHECK(d->wgt, "conv_simt: float32 weights are required");
  const int fi = d->in_fmt, fo = d->out_fmt;
  if (a.cout % 64 != 0) return launch<128, 32, 4, 4>(a, fi, fo, s);
  // enough 128x64 tiles to fill the 148 SMs at least ~2x, else use 64x64 tiles
  long long tiles128 = ((M + 127) / 128) * (a.cout / 64);
  if (tiles128 >= 296) return launch<128, 64, 8, 4>(a, fi, fo, s);
  return launch<64, 64, 4, 4>(a, fi, fo, s);
}

}  // namespace ofb
