// Training-side losses of the reference (supervision/direct.py:3-27): the reverse Huber
// (BerHu) loss of train_erp_depth_iterative.py:271 and the masked L1 loss, forward and gradient with respect to the
// prediction.  HBM-bound streaming reductions: three passes over (pred, gt, mask, weights) for the forward (the
// threshold c = max|gt - pred| / 5 is a global quantity that every element's loss depends on), one for the backward.
//
// Reference semantics reproduced on purpose:
//  * c comes from the UNMASKED maximum of |gt - pred| over the whole batch and is a constant for autograd (.item());
//    it is computed in double (a Python float) and rounded to float where it meets the float32 tensors;
//  * loss_i = |d_i| if |d_i| <= c else (d_i^2 + c^2) / (2 c);  c == 0 makes the second branch 0 / 0 = NaN and
//    0 * NaN = NaN poisons the result exactly as in torch;
//  * per sample: sum(loss * mask * weights) / sum(mask); the batch mean of those (a sample without valid pixels gives
//    0 / 0 = NaN, as in the reference).
#include "common.cuh"

namespace ofb {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_CHUNKS = 64;          // blocks per sample: partial sums are reduced in a fixed order

__global__ void __launch_bounds__(LOSS_THREADS)
loss_max_kernel(const float* __restrict__ pred, const float* __restrict__ gt, size_t n, unsigned int* __restrict__ maxbits) {
  float m = 0.f;
  bool nan = false;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a = fabsf(gt[i] - pred[i]);
    nan = nan || a != a;
    m = fmaxf(m, a);
  }
  if (nan) m = __uint_as_float(0x7FC00000u);      // torch.max propagates NaN
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, m, o);
    m = (m != m || other != other) ? __uint_as_float(0x7FC00000u) : fmaxf(m, other);
  }
  // non-negative floats (and the canonical NaN above them) order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(maxbits, __float_as_uint(m));
}

// partial[(b * LOSS_CHUNKS + chunk) * 2 + {0, 1}] = sum of loss * mask * weight, sum of mask over the chunk (double)
__global__ void __launch_bounds__(LOSS_THREADS)
loss_sum_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask,
                const float* __restrict__ weights, size_t per_sample, int berhu, const unsigned int* __restrict__ maxbits,
                double* __restrict__ partial) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const double cd = (double)__uint_as_float(*maxbits) / 5.0;
  const float c = (float)cd, c2 = (float)(cd * cd), den = (float)(2.0 * cd);
  const size_t len = (per_sample + LOSS_CHUNKS - 1) / LOSS_CHUNKS;
  const size_t i0 = (size_t)chunk * len, i1 = i0 + len < per_sample ? i0 + len : per_sample;
  const size_t base = (size_t)b * per_sample;
  double s = 0.0, cnt = 0.0;
  for (size_t i = i0 + threadIdx.x; i < i1; i += LOSS_THREADS) {
    const float d = gt[base + i] - pred[base + i];
    const float a = fabsf(d);
    float l = a;
    if (berhu) {
      const float leq = a <= c ? 1.f : 0.f;
      const float l2 = (d * d + c2) / den;
      l = leq * a + (1.f - leq) * l2;            // both terms, like the reference: 0 * NaN stays NaN
    }
    const float m = mask[base + i];
    float t = l * m;
    if (weights) t *= weights[base + i];
    s += (double)t;
    cnt += (double)m;
  }
  __shared__ double sh[2][LOSS_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tc = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) { ts += sh[0][w]; tc += sh[1][w]; }
    partial[((size_t)b * LOSS_CHUNKS + chunk) * 2] = ts;
    partial[((size_t)b * LOSS_CHUNKS + chunk) * 2 + 1] = tc;
  }
}

// loss = mean_b (sum_b / count_b); stats[0] = c, stats[1 + b] = count_b (for the backward)
__global__ void loss_final_kernel(const double* __restrict__ partial, int bs, const unsigned int* __restrict__ maxbits,
                                  float* __restrict__ loss, float* __restrict__ stats) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int b = 0; b < bs; ++b) {
    double s = 0.0, c = 0.0;
    for (int k = 0; k < LOSS_CHUNKS; ++k) { s += partial[((size_t)b * LOSS_CHUNKS + k) * 2]; c += partial[((size_t)b * LOSS_CHUNKS + k) * 2 + 1]; }
    stats[1 + b] = (float)c;
    acc += (double)((float)s / (float)c);
  }
  *loss = (float)(acc / bs);
  stats[0] = (float)((double)__uint_as_float(*maxbits) / 5.0);
}

// grad_pred_i = g * mask_i * w_i / (count_b * bs) * d loss_i / d pred_i, with
// d|d|/dpred = -sign(d), d((d^2 + c^2) / 2c)/dpred = -2 d / (2c)  (c is a constant)
__global__ void __launch_bounds__(LOSS_THREADS)
loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ mask,
                 const float* __restrict__ weights, size_t per_sample, int bs, int berhu, const float* __restrict__ stats,
                 const float* __restrict__ grad_out, float* __restrict__ grad_pred) {
  const size_t n = per_sample * bs;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = (int)(i / per_sample);
  const double cd = (double)stats[0];
  const float c = stats[0], den = (float)(2.0 * cd);
  const float d = gt[i] - pred[i];
  const float a = fabsf(d);
  const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  float dl = -sgn;
  if (berhu) {
    const float leq = a <= c ? 1.f : 0.f;
    dl = leq * (-sgn) + (1.f - leq) * ((-2.f * d) / den);
  }
  float g = *grad_out / (float)bs / stats[1 + b] * mask[i];
  if (weights) g *= weights[i];
  grad_pred[i] = g * dl;
}

}  // namespace ofb

using namespace ofb;

// work: device scratch of ofb_loss_work_bytes(bs) bytes.  stats (device, 1 + bs floats) receives c and the per-sample
// valid counts the backward needs.  mode 1 = BerHu (weights may be NULL = all ones), 0 = masked L1.
extern "C" long long ofb_loss_work_bytes(int bs) { return 16 + (long long)bs * LOSS_CHUNKS * 2 * sizeof(double); }

extern "C" int ofb_depth_loss_f32(const float* pred, const float* gt, const float* mask, const float* weights, int bs,
                                  long long per_sample, int mode, void* work, float* stats, float* loss, void* stream) {
  OFB_CHECK(pred && gt && mask && work && stats && loss && bs > 0 && per_sample > 0, "depth_loss: bad arguments");
  OFB_CHECK(mode == 0 || mode == 1, "depth_loss: mode must be 0 (L1) or 1 (BerHu), got %d", mode);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned int* maxbits = reinterpret_cast<unsigned int*>(work);
  double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(work) + 16);
  const size_t n = (size_t)bs * per_sample;
  OFB_CUDA(cudaMemsetAsync(maxbits, 0, 16, s));
  if (mode == 1) {
    const int blocks = (int)((n + LOSS_THREADS * 8 - 1) / (LOSS_THREADS * 8) < 1184 ? (n + LOSS_THREADS * 8 - 1) / (LOSS_THREADS * 8) : 1184);
    loss_max_kernel<<<blocks, LOSS_THREADS, 0, s>>>(pred, gt, n, maxbits);
    OFB_LAUNCH_CHECK();
  }
  loss_sum_kernel<<<dim3(LOSS_CHUNKS, bs), LOSS_THREADS, 0, s>>>(pred, gt, mask, weights, (size_t)per_sample, mode, maxbits, partial);
  OFB_LAUNCH_CHECK();
  loss_final_kernel<<<1, 32, 0, s>>>(partial, bs, maxbits, loss, stats);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_depth_loss_backward_f32(const float* pred, const float* gt, const float* mask, const float* weights,
                                           int bs, long long per_sample, int mode, const float* stats,
                                           const float* grad_out, float* grad_pred, void* stream) {
  OFB_CHECK(pred && gt && mask && stats && grad_out && grad_pred && bs > 0 && per_sample > 0, "depth_loss_backward: bad arguments");
  OFB_CHECK(mode == 0 || mode == 1, "depth_loss_backward: mode must be 0 (L1) or 1 (BerHu), got %d", mode);
  const size_t n = (size_t)bs * per_sample;
  loss_grad_kernel<<<(unsigned)((n + LOSS_THREADS - 1) / LOSS_THREADS), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      pred, gt, mask, weights, (size_t)per_sample, bs, mode, stats, grad_out, grad_pred);
  OFB_LAUNCH_CHECK();
  return 0;
}
