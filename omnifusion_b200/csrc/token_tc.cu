// The transformer stack of the patch network (model/blocks.py:50-88: six Transformer_Blocks over the N patch
// tokens of a panorama, width 512, 4 heads of 128, MLP 2048; spherical_model_iterative.py:330-335) as ONE launch.
//
// Attention only mixes the tokens of one panorama, so a panorama is a closed problem: a GROUP OF 16 CTAs per
// panorama (consecutive block indices; 9 groups are resident on 148 SMs) walks all blocks without any grid-wide
// synchronisation.  The work of a block is 3.1 M weights against 18-46 tokens: it is bound by streaming the weights
// out of L2, so every CTA of the group streams 1/16 of them through its own TMA ring, and the GEMMs run "swapped":
// the WEIGHT rows are the M = 128 dimension of a tcgen05 MMA, the tokens (padded to NP = 32 or 48) its N dimension.
//
//   CTA r = (head h = r / 4, quarter qd = r % 4) of its group
//   phase 0  q|k|v rows of dims [32 qd, 32 qd + 32) of head h   (3 x 32 weight rows, K = 512)  -> partial scores over
//            its 32 dims -> exchange -> softmax (replicated in the 4 CTAs of a head) -> P V for its 32 dims
//   phase 1  attn.proj rows [32 r, 32 r + 32)                    (K = 512)  + bias + residual -> exchange
//   phase 2  mlp.fc1 rows [128 r, 128 r + 128)                   (K = 512)  + bias, GELU -> stays in shared memory
//   phase 3  mlp.fc2, K slice [128 r, 128 r + 128) of all 512 rows (its own fc1 outputs are exactly that slice)
//            -> partial sums -> exchange -> fixed-order reduction of rows [32 r, 32 r + 32) + bias + residual -> exchange
//
// Operand precision is the split-half scheme of the conv engine: weights and activations are (hi, lo) fp16 pairs,
// the three products hi*hi + hi*lo + lo*hi accumulate in fp32 in TMEM.  The B operand (tokens) holds the hi rows
// and the lo rows of a 64-wide K chunk back to back, so Whi x [Xhi ; Xlo] is one MMA of N = 2 NP and Wlo x Xhi one of
// N = NP: accumulator columns [0, NP) + [NP, 2 NP) are added by the epilogue.
//
// Exchanges between the CTAs of a group go through small global (L2-resident) buffers: writers store, one thread
// per CTA adds to the panorama's arrival counter (red.release.gpu) and polls it (ld.acquire.gpu), readers use
// ld.global.cg.  (A hardware cluster of 16 with remote mbarrier arrives works too, but B200 keeps only 7 such
// clusters resident: 8 panoramas took two waves.)  A group that is only partly resident spins until the earlier
// groups retire - block dispatch is in index order, the scheme of every decoupled look-back scan.  Five exchanges
// per block; a LayerNorm is computed redundantly by every CTA from the exchanged fp32 residual stream while it
// builds its B operand.  The weight stream does not depend on any of this: a producer warp runs ahead through the
// whole stack (it starts before the previous kernel has finished), so the ring is full whenever a phase starts.
// Results do not depend on the batch size or position (one group per panorama, fixed summation orders).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tc_ptx.cuh"
#include "token_tc.cuh"

namespace ofb {

constexpr int TK_CL = 16;                 // CTAs per panorama (a "group": consecutive block indices)
constexpr int TK_PLANE = 128 * 128;       // one weight tile of one plane: 128 rows x 64 fp16
constexpr int TK_STAGE = 2 * TK_PLANE;    // ring stage: hi tile, lo tile
constexpr int TK_WORKERS = 256;           // warps 0-7; warp 8 = weight producer, warp 9 = MMA issue
constexpr int TK_THREADS = 320;

template <int NP>
struct TokCfg {
  static constexpr int CH = 2 * NP * 128;                 // bytes of one K chunk of a B operand (NP hi rows, NP lo rows)
  static constexpr int XB = 8 * CH, HB = 2 * CH;          // K = 512 operand, K = 128 operand (own fc1 outputs)
  static constexpr int QP = 33;                           // row pitch of the float32 scratch tiles
  static constexpr int F32_WORDS = 3 * NP * QP + NP * QP + NP * (NP + 1);
  static constexpr int BAR_BYTES = 512;
  static constexpr int FIXED = 1024 + XB + HB + F32_WORDS * 4 + BAR_BYTES;
  static constexpr int NST_RAW = (227 * 1024 - FIXED) / TK_STAGE;
  static constexpr int NST = NST_RAW > 4 ? 4 : NST_RAW;
  static constexpr int SMEM = FIXED + NST * TK_STAGE;
  static_assert(NP % 16 == 0 && NP <= 64 && NST >= 2 && 8 * NP <= 512, "token tile");
};

__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Waits of this kernel are bounded in TIME (a group may legitimately wait for a whole wave of earlier groups, and a
// sanitizer slows everything down by orders of magnitude): a broken hand-off faults after four seconds instead of
// hanging the GPU.
__device__ __forceinline__ bool tk_expired(int& spins, long long& t0) {
  if ((++spins & 1023) != 0) return false;
  const long long now = globaltimer_ns();
  if (t0 == 0) t0 = now;
  return now - t0 > 4000000000LL;
}
__device__ __forceinline__ void tk_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (tk_expired(spins, t0)) __trap();
  }
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TK_WORKERS) : "memory"); }
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// Rows warp, warp + 8, ... (R of them, the ones < N) of a (N, 512) float32 matrix -> B operand: hi rows and lo rows per
// 64-wide K chunk, 128-byte swizzle; optionally through LayerNorm (two passes in registers like layernorm_kernel).
// All R rows of a warp are loaded before the first is used and their reductions are interleaved.
template <int NP, int R>
__device__ __forceinline__ void fill_rows(uint8_t* xb_ptr, const float* __restrict__ src, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps, bool ln, int N, int warp, int lane) {
  constexpr int CH = TokCfg<NP>::CH;
  float v[R][16];
  float4 gm[4], bt[4];          // fetched together with the rows: their latency hides behind the same round trip
  if (ln) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gm[i] = __ldg(reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4));
      bt[i] = __ldg(reinterpret_cast<const float4*>(beta + i * 128 + lane * 4));
    }
  }
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int t = warp + 8 * u < N ? warp + 8 * u : 0;       // rows beyond N: load row 0, never stored
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (size_t)t * 512 + i * 128 + lane * 4));
      v[u][4 * i] = a.x; v[u][4 * i + 1] = a.y; v[u][4 * i + 2] = a.z; v[u][4 * i + 3] = a.w;
    }
  }
  if (ln) {
    float s[R], q[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s[u] += v[u][i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      s[u] = s[u] / 512.f;
      q[u] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = v[u][i] - s[u]; q[u] += d * d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < R; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const float var = q[u] / 512.f + eps;
      float rstd = rsqrtf(var);
      q[u] = rstd * (1.5f - 0.5f * var * rstd * rstd);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        v[u][4 * i] = (v[u][4 * i] - s[u]) * q[u] * gm[i].x + bt[i].x;
        v[u][4 * i + 1] = (v[u][4 * i + 1] - s[u]) * q[u] * gm[i].y + bt[i].y;
        v[u][4 * i + 2] = (v[u][4 * i + 2] - s[u]) * q[u] * gm[i].z + bt[i].z;
        v[u][4 * i + 3] = (v[u][4 * i + 3] - s[u]) * q[u] * gm[i].w + bt[i].w;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int t = warp + 8 * u;
    if (t < N) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = i * 128 + lane * 4;                      // four consecutive K elements = 8 bytes of one 16-byte chunk
        const uint32_t off = (uint32_t)(k >> 6) * CH + (uint32_t)t * 128 +
                             (((uint32_t)((k & 63) >> 3) ^ ((uint32_t)t & 7)) << 4) + (uint32_t)(k & 7) * 2;
        const __half2 h0 = __floats2half2_rn(v[u][4 * i], v[u][4 * i + 1]), h1 = __floats2half2_rn(v[u][4 * i + 2], v[u][4 * i + 3]);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(v[u][4 * i] - f0.x, v[u][4 * i + 1] - f0.y);
        const __half2 l1 = __floats2half2_rn(v[u][4 * i + 2] - f1.x, v[u][4 * i + 3] - f1.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
        lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(xb_ptr + off) = hv;
        *reinterpret_cast<uint2*>(xb_ptr + off + NP * 128) = lv;      // lo rows: NP rows further (NP % 8 == 0: same swizzle phase)
      }
    }
  }
}

// this warp's half of the accumulator columns of its 32 TMEM lanes: hi-product columns + lo-product columns
template <int NP>
__device__ __forceinline__ void acc_load(uint32_t taddr, int c_begin, float* out) {
  constexpr int W = NP / 2;
  uint32_t a[W], l[W];
#pragma unroll
  for (int c = 0; c < W; c += 8) {
    tmem_ld8_issue(taddr + c_begin + c, a + c);
    tmem_ld8_issue(taddr + NP + c_begin + c, l + c);
  }
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < W; ++c) out[c] = __uint_as_float(a[c]) + __uint_as_float(l[c]);
}

template <int NP>
__global__ void __launch_bounds__(TK_THREADS, 1)
token_stack_kernel(const __grid_constant__ TokStack st, float* __restrict__ xg, float* __restrict__ sg,
                   float* __restrict__ ag, float* __restrict__ pg, float* __restrict__ enc, int N, int stop,
                   unsigned int* __restrict__ counters, long long* __restrict__ stamps) {
  using C = TokCfg<NP>;
  constexpr int NST = C::NST, CH = C::CH, QP = C::QP, W = NP / 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023) & ~1023u;
  uint8_t* bp = smem_raw + (base - raw);
  const uint32_t xb = base, hb = base + C::XB, ring = hb + C::HB;
  uint8_t* xb_ptr = bp;
  uint8_t* hb_ptr = bp + C::XB;
  float* qs = reinterpret_cast<float*>(bp + C::XB + C::HB + NST * TK_STAGE);      // [3: q k v][NP tokens][QP]
  float* xres = qs + 3 * NP * QP;                                                    // [NP][QP] residual stream, columns 32 r ..
  float* ps = xres + NP * QP;                                                        // [NP][NP + 1] scores / probabilities
  const uint32_t bars = ring + NST * TK_STAGE + C::F32_WORDS * 4;
  const uint32_t bar_full = bars, bar_empty = bars + 64, bar_ready = bars + 128, bar_acc = bars + 136;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bp + C::XB + C::HB + NST * TK_STAGE + C::F32_WORDS * 4 + 160);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = blockIdx.x % TK_CL;
  const int pano = blockIdx.x / TK_CL;
  const int head = rank >> 2, qd = rank & 3;
  const int nph = stop > 0 ? stop : 4 * st.nblk;            // GEMM phases to run

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_ready, 1);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 8) {
    // ---------------------------------------------------------------- weight producer
    // The stream is a fixed sequence: per phase eight (row tile, K chunk) tiles, each the hi and the lo plane.  A tensor
    // map covers [hi plane ; lo plane] of one linear as (2 * cout) rows of cin halves, box = 64 halves x 32 or 128 rows.
    uint32_t it = 0;
    for (int g = 0; g < nph; ++g) {
      const int ph = g & 3;
      const CUtensorMap* map = &st.blk[g >> 2].w[ph];
#pragma unroll 1
      for (int tile = 0; tile < 8; ++tile, ++it) {
        const uint32_t slot = it % NST, use = it / NST;
        if (use > 0) tk_wait(bar_empty + 8 * slot, (use - 1) & 1);
        if (elect_one()) {
          const uint32_t dst = ring + slot * TK_STAGE, full = bar_full + 8 * slot;
          if (ph == 0) {
            mbar_expect_tx(full, 6 * 4096);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
              for (int s3 = 0; s3 < 3; ++s3)
                tma_load_2d(dst + pl * TK_PLANE + s3 * 4096, map, full, tile * 64, pl * 1536 + s3 * 512 + 128 * head + 32 * qd);
          } else if (ph == 1) {
            mbar_expect_tx(full, 2 * 4096);
            tma_load_2d(dst, map, full, tile * 64, 32 * rank);
            tma_load_2d(dst + TK_PLANE, map, full, tile * 64, 512 + 32 * rank);
          } else if (ph == 2) {
            mbar_expect_tx(full, 2 * 16384);
            tma_load_2d(dst, map, full, tile * 64, 128 * rank);
            tma_load_2d(dst + TK_PLANE, map, full, tile * 64, 2048 + 128 * rank);
          } else {
            mbar_expect_tx(full, 2 * 16384);
            tma_load_2d(dst, map, full, 128 * rank + 64 * (tile & 1), 128 * (tile >> 1));
            tma_load_2d(dst + TK_PLANE, map, full, 128 * rank + 64 * (tile & 1), 512 + 128 * (tile >> 1));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ---------------------------------------------------------------- MMA issue
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)((2 * NP) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t dconst = umma_desc<128>(0);
    uint32_t it = 0;
    for (int g = 0; g < nph; ++g) {
      const int ph = g & 3;
      tk_wait(bar_ready, g & 1);
      tc_fence_after();
      const uint32_t bb = ph == 3 ? hb : xb;
#pragma unroll 1
      for (int tile = 0; tile < 8; ++tile, ++it) {
        const int chunk = ph == 3 ? (tile & 1) : tile;
        const uint32_t dcol = ph == 3 ? (uint32_t)(tile >> 1) * 2 * NP : 0u;
        const uint32_t first = (ph == 3 ? (tile & 1) == 0 : tile == 0) ? 0u : 1u;
        const uint64_t bdesc = dconst | (uint64_t)(((bb + chunk * CH) >> 4) & 0x3FFF);
        const uint32_t slot = it % NST;
        tk_wait(bar_full + 8 * slot, (it / NST) & 1);
        tc_fence_after();
        const uint64_t ahi = dconst | (uint64_t)(((ring + slot * TK_STAGE) >> 4) & 0x3FFF);
        const uint64_t alo = dconst | (uint64_t)(((ring + slot * TK_STAGE + TK_PLANE) >> 4) & 0x3FFF);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {      // Whi x [Xhi ; Xlo] then Wlo x Xhi per K16 slice
            tc_mma<MODE_F16X3>(tmem + dcol, ahi + 2 * kk, bdesc + 2 * kk, idesc1, kk == 0 ? first : 1u);
            tc_mma<MODE_F16X3>(tmem + dcol, alo + 2 * kk, bdesc + 2 * kk, idesc2, 1u);
          }
          tc_commit(bar_empty + 8 * slot);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(bar_acc);
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- workers
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int q4 = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem + ((uint32_t)(q4 * 32) << 16);
    const size_t xoff = (size_t)pano * N * 512;
    const int rpw = (N + 7) >> 3;
    uint32_t apar = 0;
    unsigned int xtarget = 0;
    unsigned int* ctr = counters + 2 * pano;
    for (int i = tid; i < N * 32; i += TK_WORKERS) xres[(i >> 5) * QP + (i & 31)] = __ldcg(xg + xoff + (size_t)(i >> 5) * 512 + 32 * rank + (i & 31));

    auto fill_xb = [&](const float* src, const float* gamma, const float* beta, float eps, bool ln) {
      if (rpw <= 2) fill_rows<NP, 2>(xb_ptr, src, gamma, beta, eps, ln, N, warp, lane);
      else if (rpw == 3) fill_rows<NP, 3>(xb_ptr, src, gamma, beta, eps, ln, N, warp, lane);
      else if (rpw == 4 || NP == 32) fill_rows<NP, 4>(xb_ptr, src, gamma, beta, eps, ln, N, warp, lane);
      else fill_rows<NP, NP / 8>(xb_ptr, src, gamma, beta, eps, ln, N, warp, lane);
    };
    // operand complete -> MMA warp
    auto operand_ready = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      worker_sync();
      if (tid == 0) mbar_arrive(bar_ready);
    };
    auto wait_acc = [&]() {
      tk_wait(bar_acc, apar);
      apar ^= 1;
      tc_fence_after();
    };
    // all 16 CTAs of the panorama have stored what they exchange -> all may read it (ld.global.cg).  One arrival
    // counter per panorama in global memory: release-add by one thread per CTA, acquire-poll by the same thread.
    auto exchange = [&]() {
      xtarget += TK_CL;
      worker_sync();
      if (tid == 0) {
        red_release_gpu(ctr, 1u);
        int spins = 0;
        long long t0 = 0;
        while (ld_acquire_gpu(ctr) < xtarget) { if (tk_expired(spins, t0)) __trap(); }
      }
      worker_sync();
    };
    // experiments: clock stamps of thread 0 of CTA 0 (tools/probe_token.py): [phase][8]
    const bool stamping = stamps != nullptr && blockIdx.x == 0 && tid == 0;
    auto stamp = [&](int g, int k) { if (stamping) stamps[g * 8 + k] = clock64(); };

    for (int g = 0; g < nph; ++g) {
      const int ph = g & 3;
      const TokBlock& B = st.blk[g >> 2];
      stamp(g, 0);
      if (ph == 0) {
        fill_xb(xg + xoff, B.n1g, B.n1b, 1e-5f, true);
        stamp(g, 1);
        operand_ready();
        wait_acc();
        stamp(g, 2);
        if (q4 < 3) {          // TMEM lanes 0-31 q, 32-63 k, 64-95 v rows of this CTA's 32 dims; lanes 96-127 are unused
          const int frow = q4 * 512 + 128 * head + 32 * qd + lane;
          const float bias = B.bias[0] ? __ldg(B.bias[0] + frow) : 0.f, us = B.unscale[0];   // q / kv projections have no bias
          float acc[W];
          acc_load<NP>(trow, half * W, acc);
#pragma unroll
          for (int j = 0; j < W; ++j) qs[(q4 * NP + half * W + j) * QP + lane] = acc[j] * us + bias;
        }
        tc_fence_before();
        worker_sync();
        // partial scores over this CTA's 32 dims
        const int NN = N * N;
        float* sdst = sg + ((size_t)(pano * 4 + head) * 4 + qd) * NN;
        for (int p = tid; p < NN; p += TK_WORKERS) {
          const int i = p / N, j = p - i * N;
          const float* qr = qs + i * QP;
          const float* kr = qs + (NP + j) * QP;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int d = 0; d < 32; d += 4) {
            a0 += qr[d] * kr[d]; a1 += qr[d + 1] * kr[d + 1]; a2 += qr[d + 2] * kr[d + 2]; a3 += qr[d + 3] * kr[d + 3];
          }
          sdst[p] = (a0 + a1) + (a2 + a3);
        }
        stamp(g, 3);
        exchange();
        stamp(g, 4);
        // scores of the head = the four partial sums in quarter order; softmax with 8 lanes per query row
        const float* ssrc = sg + (size_t)(pano * 4 + head) * 4 * NN;
        const float scale = 0.08838834764831845f;       // 1 / sqrt(128)
        for (int p0 = tid; p0 < NN; p0 += 2 * TK_WORKERS) {
          const int p1 = p0 + TK_WORKERS;
          float s0[4], s1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            s0[u] = __ldcg(ssrc + u * NN + p0);
            s1[u] = p1 < NN ? __ldcg(ssrc + u * NN + p1) : 0.f;
          }
          const int i0 = p0 / N;
          ps[i0 * (NP + 1) + (p0 - i0 * N)] = (((s0[0] + s0[1]) + s0[2]) + s0[3]) * scale;
          if (p1 < NN) {
            const int i1 = p1 / N;
            ps[i1 * (NP + 1) + (p1 - i1 * N)] = (((s1[0] + s1[1]) + s1[2]) + s1[3]) * scale;
          }
        }
        worker_sync();
        stamp(g, 7);
        for (int i0 = warp * 4; i0 < N; i0 += TK_WORKERS / 8) {      // four rows per warp: the loop is warp-uniform
          const int i = i0 + (lane >> 3), l8 = lane & 7;
          const bool row = i < N;
          float* pr = ps + (row ? i : 0) * (NP + 1);
          float m = -INFINITY;
          if (row) for (int j = l8; j < N; j += 8) m = fmaxf(m, pr[j]);
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          float s = 0.f;
          if (row) for (int j = l8; j < N; j += 8) { const float e = expf(pr[j] - m); pr[j] = e; s += e; }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float inv = 1.f / s;
          if (row) for (int j = l8; j < N; j += 8) pr[j] *= inv;
        }
        worker_sync();
        for (int o = tid; o < N * 32; o += TK_WORKERS) {
          const int i = o >> 5, d = o & 31;
          const float* pr = ps + i * (NP + 1);
          float a0 = 0.f, a1 = 0.f;          // even / odd keys, ascending
#pragma unroll 1
          for (int j0 = 0; j0 < N; j0 += 8) {      // eight keys per step: all sixteen shared-memory loads first
            float pj[8], vj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const bool in = j0 + u < N;
              pj[u] = in ? pr[j0 + u] : 0.f;
              vj[u] = in ? qs[(2 * NP + j0 + u) * QP + d] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) { a0 += pj[u] * vj[u]; a1 += pj[u + 1] * vj[u + 1]; }
          }
          ag[xoff + (size_t)i * 512 + 32 * rank + d] = a0 + a1;
        }
        stamp(g, 5);
        exchange();
        stamp(g, 6);
      } else if (ph == 1) {
        fill_xb(ag + xoff, nullptr, nullptr, 0.f, false);
        stamp(g, 1);
        operand_ready();
        wait_acc();
        stamp(g, 2);
        if (q4 == 0) {          // rows 0-31 of the tile = proj rows 32 r ..; y = x + proj(att) + bias
          const int frow = 32 * rank + lane;
          const float bias = __ldg(B.bias[1] + frow), us = B.unscale[1];
          float acc[W];
          acc_load<NP>(trow, half * W, acc);
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const int t = half * W + j;
            if (t < N) {
              const float y = (acc[j] * us + bias) + xres[t * QP + lane];
              xres[t * QP + lane] = y;
              xg[xoff + (size_t)t * 512 + frow] = y;
            }
          }
        }
        tc_fence_before();
        stamp(g, 3);
        exchange();
        stamp(g, 4);
      } else if (ph == 2) {
        fill_xb(xg + xoff, B.n2g, B.n2b, 1e-5f, true);
        stamp(g, 1);
        operand_ready();
        wait_acc();
        stamp(g, 2);
        {                       // h = gelu(fc1 + bias) for hidden unit 128 r + k, written as this CTA's K slice of the fc2 operand
          const int k = q4 * 32 + lane;
          const float bias = __ldg(B.bias[2] + 128 * rank + k), us = B.unscale[2];
          const uint32_t kbase = (uint32_t)(k >> 6) * CH + (uint32_t)(k & 7) * 2;
          const uint32_t kc16 = (uint32_t)(k & 63) >> 3;
          float acc[W];
          acc_load<NP>(trow, half * W, acc);
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const uint32_t t = (uint32_t)(half * W + j);
            const float hv = gelu_erf(acc[j] * us + bias);
            const __half hh = __float2half_rn(hv);
            const __half hl = __float2half_rn(hv - __half2float(hh));
            const uint32_t off = kbase + t * 128 + ((kc16 ^ (t & 7)) << 4);
            *reinterpret_cast<__half*>(hb_ptr + off) = hh;
            *reinterpret_cast<__half*>(hb_ptr + off + NP * 128) = hl;
          }
        }
        stamp(g, 3);
      } else {
        stamp(g, 1);
        operand_ready();
        wait_acc();
        stamp(g, 2);
        {                       // raw partial sums of all 512 fc2 rows over this CTA's K slice
          float* pdst = pg + ((size_t)(pano * TK_CL + rank) * N) * 512;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int frow = 128 * m + 32 * q4 + lane;
            float acc[W];
            acc_load<NP>(trow + m * 2 * NP, half * W, acc);
#pragma unroll
            for (int j = 0; j < W; ++j) {
              const int t = half * W + j;
              if (t < N) pdst[(size_t)t * 512 + frow] = acc[j];
            }
          }
        }
        tc_fence_before();
        stamp(g, 3);
        exchange();
        stamp(g, 4);
        {                       // x = y + fc2 + bias for columns 32 r ..: the 16 partial sums in rank order
          const float us = B.unscale[3];
          for (int o = tid; o < N * 32; o += TK_WORKERS) {
            const int t = o >> 5, l = o & 31;
            const int f = 32 * rank + l;
            const float* src = pg + ((size_t)pano * TK_CL * N + t) * 512 + f;
            float v[TK_CL];
#pragma unroll
            for (int s = 0; s < TK_CL; ++s) v[s] = __ldcg(src + (size_t)s * N * 512);
            float acc = 0.f;
#pragma unroll
            for (int s = 0; s < TK_CL; ++s) acc += v[s];
            const float y = (acc * us + __ldg(B.bias[3] + f)) + xres[t * QP + l];
            xres[t * QP + l] = y;
            xg[xoff + (size_t)t * 512 + f] = y;
          }
        }
        stamp(g, 5);
        exchange();
        stamp(g, 6);
      }
    }
    if (stop == 0) {
      // encoder_norm (eps 1e-6): the panorama's rows are spread over its 16 CTAs
      const int t = rank + TK_CL * warp;
      if (t < N) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(xg + xoff + (size_t)t * 512 + i * 128 + lane * 4));
          v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / 512.f;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float var = q / 512.f + 1e-6f;
        float rstd = rsqrtf(var);
        rstd = rstd * (1.5f - 0.5f * var * rstd * rstd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(st.enc_g + i * 128 + lane * 4));
          const float4 bt = __ldg(reinterpret_cast<const float4*>(st.enc_b + i * 128 + lane * 4));
          float4 o;
          o.x = (v[4 * i] - mean) * rstd * gm.x + bt.x;
          o.y = (v[4 * i + 1] - mean) * rstd * gm.y + bt.y;
          o.z = (v[4 * i + 2] - mean) * rstd * gm.z + bt.z;
          o.w = (v[4 * i + 3] - mean) * rstd * gm.w + bt.w;
          *reinterpret_cast<float4*>(enc + xoff + (size_t)t * 512 + i * 128 + lane * 4) = o;
        }
      }
    }
    // the last CTA of the panorama to get here puts the two counters back to zero for the next launch (every CTA
    // has left its last poll by the time it adds to the second counter)
    worker_sync();
    if (tid == 0) {
      const unsigned int done = atomicAdd(ctr + 1, 1u);
      if (done == TK_CL - 1) { ctr[0] = 0u; ctr[1] = 0u; __threadfence(); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host
constexpr int kMaxDevices = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
static int token_np(int N) { return N <= 32 ? 32 : 48; }

bool token_stack_supported(int N) { return N >= 1 && N <= 48; }

size_t token_stack_scratch_floats(int B, int N) {
  return (size_t)B * 16 * N * N + (size_t)B * N * 512 + (size_t)B * TK_CL * N * 512;
}

void token_stack_release(TokStack* st) {
  if (st && st->counters) { cudaFree(st->counters); st->counters = nullptr; }
}

int token_stack_prepare(const TokBlockDesc* blocks, int nblk, const float* enc_g, const float* enc_b, TokStack* out) {
  OFB_CHECK(blocks && out && nblk >= 0 && nblk <= kTokMaxBlocks, "token_stack: bad arguments");
  unsigned int* keep = out->counters;
  memset(out, 0, sizeof(*out));
  out->counters = keep;
  static const int cout[4] = {1536, 512, 2048, 512}, cin[4] = {512, 512, 512, 2048}, rows[4] = {32, 32, 128, 128};
  for (int b = 0; b < nblk; ++b) {
    TokBlock& T = out->blk[b];
    for (int i = 0; i < 4; ++i) {
      const TokLinear& L = blocks[b].lin[i];
      OFB_CHECK(L.ws && (L.bias || i == 0), "token_stack: block %d linear %d has no split-half weights / bias", b, i);
      cuuint64_t dims[2] = {(cuuint64_t)cin[i], (cuuint64_t)2 * cout[i]};
      cuuint32_t box[2] = {64u, (cuuint32_t)rows[i]};
      if (tc_make_map(&T.w[i], true, 2, const_cast<void*>(L.ws), dims, box, 128)) return -1;
      T.bias[i] = L.bias;
      T.unscale[i] = L.unscale;
    }
    T.n1g = blocks[b].n1g; T.n1b = blocks[b].n1b; T.n2g = blocks[b].n2g; T.n2b = blocks[b].n2b;
    OFB_CHECK(T.n1g && T.n1b && T.n2g && T.n2b, "token_stack: block %d has no LayerNorm parameters", b);
  }
  out->enc_g = enc_g; out->enc_b = enc_b; out->nblk = nblk;
  if (!out->counters) {        // arrival counters: two per panorama and lane, zero between launches (the kernel resets them)
    OFB_CUDA(cudaMalloc(&out->counters, (size_t)kTokLanes * 2 * kTokMaxPanos * sizeof(unsigned int)));
    OFB_CUDA(cudaMemset(out->counters, 0, (size_t)kTokLanes * 2 * kTokMaxPanos * sizeof(unsigned int)));
  }
  return 0;
}

static long long* g_tok_stamps = nullptr;      // experiments (ofb_debug_token_stamps)
void token_stack_debug_stamps(long long* p) { g_tok_stamps = p; }

template <int NP>
static int token_launch(const TokStack& st, float* x, float* sg, float* ag, float* pg, float* enc, int B, int N, int stop,
                        unsigned int* counters, bool pdl, cudaStream_t s) {
  using C = TokCfg<NP>;
  static bool attr[kMaxDevices] = {false};
  const int dev = cur_device();
  if (!attr[dev]) {
    OFB_CUDA(cudaFuncSetAttribute(token_stack_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B * TK_CL); cfg.blockDim = dim3(TK_THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = s;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  OFB_CUDA(cudaLaunchKernelEx(&cfg, token_stack_kernel<NP>, st, x, sg, ag, pg, enc, N, stop, counters, g_tok_stamps));
  OFB_LAUNCH_CHECK();
  return 0;
}

// panoramas (groups of 16 CTAs) the device runs at once: one CTA per SM
int token_stack_resident_groups(int N) {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaError_t e;
  if (token_np(N) == 32) {
    cudaFuncSetAttribute(token_stack_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TokCfg<32>::SMEM);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, token_stack_kernel<32>, TK_THREADS, TokCfg<32>::SMEM);
  } else {
    cudaFuncSetAttribute(token_stack_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, TokCfg<48>::SMEM);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, token_stack_kernel<48>, TK_THREADS, TokCfg<48>::SMEM);
  }
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return sms * per_sm / TK_CL;
}

int token_stack_launch(const TokStack& st, float* x, float* scratch, size_t scratch_floats, float* enc_out, int B, int N,
                       int stop_phase, int lane, bool pdl, cudaStream_t s) {
  OFB_CHECK(x && scratch && B > 0 && token_stack_supported(N), "token_stack: bad arguments (B %d, N %d)", B, N);
  OFB_CHECK(B <= kTokMaxPanos && lane >= 0 && lane < kTokLanes && st.counters, "token_stack: at most %d panoramas per launch (got %d)", kTokMaxPanos, B);
  OFB_CHECK(stop_phase > 0 || enc_out, "token_stack: enc_out is required when the whole stack runs");
  OFB_CHECK(scratch_floats >= token_stack_scratch_floats(B, N), "token_stack: scratch holds %zu floats, %zu needed",
            scratch_floats, token_stack_scratch_floats(B, N));
  OFB_CHECK(stop_phase >= 0 && stop_phase <= 4 * st.nblk, "token_stack: stop_phase %d of %d", stop_phase, 4 * st.nblk);
  float* sg = scratch;
  float* ag = sg + (size_t)B * 16 * N * N;
  float* pg = ag + (size_t)B * N * 512;
  unsigned int* ctr = st.counters + (size_t)lane * 2 * kTokMaxPanos;
  if (token_np(N) == 32) return token_launch<32>(st, x, sg, ag, pg, enc_out, B, N, stop_phase, ctr, pdl, s);
  return token_launch<48>(st, x, sg, ag, pg, enc_out, B, N, stop_phase, ctr, pdl, s);
}

}  // namespace ofb
