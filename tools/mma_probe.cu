// Micro-probes for the tcgen05 conv engine (not part of the product):
//   1. issue/pipe time of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16, operands in shared memory) versus N;
//   2. whether a K-major swizzled A descriptor may start at a row that is NOT a multiple of 8
//      (needed for "flat shift" convolution taps), with and without the descriptor's base-offset field.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/mma_probe tools/mma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t base_off = 0) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---------------------------------------------------------------- 1. MMA rate
// mode 0: n_mma MMAs of width N; mode 1: alternating (N, N/2) pairs as the split-half conv issues them
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n_mma, int mode, int a_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  // A: 4 stages x 192 rows x 128 B; B: 4 stages x 256 rows x 128 B
  const uint32_t a0 = base, b0 = base + 4 * 192 * 128;
  for (int i = threadIdx.x; i < (4 * 192 * 128 + 4 * 256 * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(bp)[i] = 0x2c002c00u + (i & 0xff);   // small fp16 values
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1) {
    const uint32_t id_n = make_idesc(N), id_h = make_idesc(N / 2 < 8 ? 8 : N / 2);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int stg = j >> 1, kk = (j & 1) * 2 + ((i >> 3) & 1);
          // a_rows >= 0: every MMA's A tile starts a_rows * (j % 3) rows into the stage (row-shifted descriptors)
          const uint64_t ad = umma_desc<128>(a0 + stg * 192 * 128 + (uint32_t)((j % 3) * a_rows * 128)) + 2 * kk;
          const uint64_t bd = umma_desc<128>(b0 + stg * 256 * 128) + 2 * kk;
          if (mode == 1 && (j & 1)) tc_mma(tmem + 256, ad, bd, id_h, 1);
          else tc_mma(tmem, ad, bd, id_n, 1);
        }
      }
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------- 2. row-shifted A descriptors
// A: 160 rows x ROW_BYTES (swizzled by absolute address like TMA does), value A[r][k] = r + k/64 (exact in fp16 for small r).
// B: identity (KC x KC).  D[m][n] = A[m + shift][n].
template <int ROW_BYTES>
__global__ void __launch_bounds__(128, 1) shift_kernel(int shift, int use_base_off, float* out) {
  constexpr int KC = ROW_BYTES / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t a0 = base, b0 = base + 160 * ROW_BYTES + ((160 * ROW_BYTES) % 1024 ? 1024 - (160 * ROW_BYTES) % 1024 : 0);
  uint8_t* ap = bp;
  uint8_t* bq = bp + (b0 - base);
  // swizzle: 16-byte chunk index ^= (row & 7) for 128-byte rows, ^= ((row >> 1) & 3) for 64-byte rows (address based)
  for (int i = threadIdx.x; i < 160 * KC; i += blockDim.x) {
    const int r = i / KC, k = i % KC;
    const int chunk = k / 8, within = k % 8;
    const int sw = ROW_BYTES == 128 ? (chunk ^ (r & 7)) : (chunk ^ ((r >> 1) & 3));
    reinterpret_cast<__half*>(ap + r * ROW_BYTES)[sw * 8 + within] = __float2half((float)r + (float)k / 64.f);
  }
  for (int i = threadIdx.x; i < KC * KC; i += blockDim.x) {
    const int r = i / KC, k = i % KC;
    const int chunk = k / 8, within = k % 8;
    const int sw = ROW_BYTES == 128 ? (chunk ^ (r & 7)) : (chunk ^ ((r >> 1) & 3));
    reinterpret_cast<__half*>(bq + r * ROW_BYTES)[sw * 8 + within] = __float2half(r == k ? 1.f : 0.f);
  }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t astart = a0 + shift * ROW_BYTES;
      const uint32_t boff = use_base_off ? (ROW_BYTES == 128 ? (astart >> 7) & 7 : (astart >> 7) & 3) : 0;
      for (int kk = 0; kk < KC / 16; ++kk)
        tc_mma(tmem, umma_desc<ROW_BYTES>(astart, boff) + 2 * kk, umma_desc<ROW_BYTES>(b0) + 2 * kk, make_idesc(KC), kk != 0);
      tc_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int q = warp & 3, lane = threadIdx.x & 31, r = q * 32 + lane;
    for (int cb = 0; cb < KC; cb += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb, v);
      for (int j = 0; j < 32; ++j) out[r * KC + cb + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
  }
}

template <int ROW_BYTES>
static void run_shift() {
  constexpr int KC = ROW_BYTES / 2;
  float* d_out;
  CK(cudaMalloc(&d_out, 128 * KC * sizeof(float)));
  std::vector<float> h(128 * KC);
  const int smem = 160 * ROW_BYTES + 2048 + KC * ROW_BYTES + 1024;
  CK(cudaFuncSetAttribute(shift_kernel<ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int ub = 0; ub < 2; ++ub) {
    printf("row-shift ROW_BYTES=%d base_offset_field=%d: ", ROW_BYTES, ub);
    for (int shift = 0; shift <= 11; ++shift) {
      CK(cudaMemset(d_out, 0, 128 * KC * sizeof(float)));
      shift_kernel<ROW_BYTES><<<1, 128, smem>>>(shift, ub, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d: %s\n", shift, cudaGetErrorString(e)); exit(1); }
      CK(cudaMemcpy(h.data(), d_out, 128 * KC * sizeof(float), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < KC; ++n) {
          const float want = (float)(m + shift) + (float)n / 64.f;
          const float w16 = __half2float(__float2half(want));
          if (h[m * KC + n] != w16) ++bad;
        }
      printf("s%d:%s ", shift, bad ? "BAD" : "ok");
      if (bad && shift == 1 && ub == 0) printf("(e.g. D[0][0]=%g D[0][8]=%g D[1][0]=%g) ", h[0], h[8], h[KC]);
    }
    printf("\n");
  }
  CK(cudaFree(d_out));
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  const int smem = 4 * 192 * 128 + 4 * 256 * 128 + 2048;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_mma = 2048;
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {16, 32, 64, 128, 256}) {
        if (mode == 1 && N == 16) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<grid, 128, smem>>>(N, n_mma, mode, 32, d_out);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
        printf("grid=%3d mode=%d N=%3d%s: issue %.1f clk/MMA, complete %.1f clk/MMA\n", grid, mode, N,
               mode ? " (alternating N, N/2)" : "", (double)h[0] / n_mma, (double)h[1] / n_mma);
      }
  }
  for (int sh : {0, 1, 2, 3, 4, 8, 16, 32})
    for (int N : {16, 32, 64}) {
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<<<148, 128, smem>>>(N, n_mma, 0, sh, d_out);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
      printf("row shift %2d x (0,1,2) N=%3d: %.1f clk/MMA\n", sh, N, (double)h[1] / n_mma);
    }
  run_shift<128>();
  run_shift<64>();
  return 0;
}
