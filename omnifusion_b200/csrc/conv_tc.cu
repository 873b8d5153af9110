// tcgen05 / TMEM / TMA implicit-GEMM convolution engine (sm_100a): every 3x3 / 1x1 conv (stride 1 and 2), every
// nn.Linear, the 7x7 stem and the two heads of the OmniFusion network.
//
// GEMM view: M = output pixels (CTA tile = 128 pixels = a TMA box of BNI images x BH rows x BW columns), N = cout
// (CTA tile BN <= 128), K = taps x channels.  For every filter tap the producer warp issues one 4-D TMA load of
// the *shifted* activation box (out-of-bounds rows/columns are zero-filled by TMA = the conv padding; traversal
// strides = the conv stride) and one load of the matching K-slice of the OHWI weights; both land in shared memory
// in the K-major 128B/64B-swizzled layout that tcgen05.mma consumes directly.  The MMA warp accumulates in TMEM
// (two accumulator buffers: tile i+1 runs while the epilogue drains tile i); epilogue warps read TMEM back
// (tcgen05.ld), apply scale/shift/residual/activation, convert to the storage format and store through staged
// bulk tensor stores.  Persistent CTAs, NST-stage smem ring with full/empty mbarriers, tcgen05.commit frees stages.
// The producer and MMA warps run warp-uniformly and predicate only the instruction issue on elect.sync.
//
// Operand modes:
//   TF32   : float32 tensors, one kind::tf32 MMA per K-step (TF32 accuracy; cross-check only).
//   F16X3  : split-half planes (value = hi + lo, both fp16): hi * [Whi; Wlo] as ONE MMA against the stacked weight
//            tile plus lo * Whi, accumulated in fp32 - ~22-bit operands, fp32-level result.
// Variants (template flags, see TcCfg): KHR kh-reuse boxes for narrow 3x3 layers, BRES resident filter, UPS rolling
// rows (1: fused 2x upsample with interpolating producer warps, 2: TMA-fed rows + heads epilogue), CTA2
// cta_group::2 CTA pairs for the 128-wide tiles; split-K for the token linears (TcParams::ksplit).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tc_ptx.cuh"

namespace ofb {


struct TcParams {
  int n_img, H, W, c0, c1, cout, k, pad, stride;   // H, W: output dims
  // tap -> TMA coordinates: x = x0*sx + kw - padx, y = y0*sy + kh - pady with kh = tap / kdiv, kw = tap % kdiv
  int taps, kdiv, sx, sy, padx, pady;
  int BW, BH, BNI, tiles_x, tiles_y;
  const float* scale; const float* shift; float wscale;
  const void* residual; void* out;
  int act;
  long long plane;   // elements per plane of out / residual (split format)
  int tiles_n, total_tiles;
  const void* ups_src;   // UPS: low-resolution input (n_img, H/2, W/2, c0), split-half planes
  long long ups_plane;   // elements per plane of ups_src
  int dbg;               // timing experiments only (results are wrong):
                         // 32 skip the fused-upsample interpolation, 64 skip the epilogue math and stores,
                         // 128 skip only the global stores of the epilogue
  int ksplit;            // split-K (1x1 / linear layers only): tile index = ((m tile * tiles_n + n tile) * ksplit + split);
  float* partial;        // each split writes its raw float32 accumulators to partial[split][pixel][cout] (ofb_splitk_finish_ln_f32 reduces)
  long long m_total;     // pixels per split plane of `partial`
  float* heads_pred; float* heads_conf; float heads_bp, heads_bc;   // UPS == 2: the two heads' outputs and biases
  int heads_il;          // UPS == 2: write (pred * conf, conf) as ONE float2 per pixel into heads_pred (the layout the
                         // blend kernel gathers with one 8-byte load per tap) instead of two separate maps
  long long* dbg_buf;    // dbg & 16: clock stamps of the epilogue warp 2 of CTA 0: [tile][8]
                         // dbg & 256: %globaltimer stamps of CTA 0 of every launch: [launch slot][8] (kTimelineSlots slots,
                         // slot counter behind them): 0 kernel entry, 1 prologue done, 2 producer past griddepcontrol.wait,
                         // 3 first operands landed, 4 last MMA issued, 5 first accumulator complete, 6 last store issued,
                         // 7 kernel end
  int group64;           // accumulate the three split-half products in the cta_group::2 column grouping (see the MMA warp)
  int nstack;            // UPS kernels: input-row-stationary MMAs with the three kh taps stacked along N (see the MMA warp)
  int wmc;               // kh-reuse kernels (weights in the ring): launched as clusters of two CTAs that walk their tiles
                         // in lockstep; each loads HALF of a stage's weight box and multicasts it to both
};

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == OFB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == OFB_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  return v;
}

// Per-configuration constants.  TMA_STORE: the epilogue stages 32-column chunks in shared
// memory (swizzled) and writes them with cp.async.bulk.tensor stores (coalesced, asynchronous);
// used for the narrow-N, HBM-bound layers.  Wide tiles store straight from registers.
// KHR ("kh reuse", 3x3 stride-1 convs whose tile lies inside one image): one stage holds the
// activation box with a one-row halo above and below ((BH+2) x BW pixels, loaded once per kw) and
// the weight slices of all three kh taps; the three kh MMAs read the same box at row offsets
// 0, BW, 2*BW (multiples of 8 rows, so the swizzle phase is unchanged).  A-operand traffic from
// L2 drops from 9*BH to 3*(BH+2) image rows per tile - the fix for the L2-bound narrow layers.
// BRES (KHR layers whose whole filter is a few KB, i.e. 32->32): the weights are loaded once per CTA
// into a resident region and the ring stages carry activations only - the layer is bound by the TMA
// request rate on its 64-byte rows, and a third of those requests were weight re-loads.
// UPS (BRES layers only, image width 128): the conv input is the 2x bilinear upsample (align_corners=False) of a
// low-resolution tensor, which never exists in HBM.  A CTA walks a contiguous range of output rows (tile =
// one image row of 128 pixels) and keeps a ring of UPSAMPLED rows in shared memory: every output row adds ONE
// new row (130 pixels: zero column, 128 interpolated pixels, zero column) that eight producer warps
// interpolate from the low-resolution tensor (thread = low-res column x 8-channel chunk; the horizontally
// blended low-res rows are cached in registers, so a row costs 16 lerps and 4 shared stores per thread).
// All nine taps read the ring through ROW-SHIFTED operand descriptors: tap (kh, kw) of output row y is the 128
// consecutive ring rows starting at pixel kw of upsampled row y-1+kh (the swizzle is a function of the
// shared-memory address, so a descriptor may start at any 64-byte row - verified by tools/mma_probe.cu).
// Against the tile-box scheme (three kw-shifted copies of a 6-row halo box per 4x32 tile) the producers
// interpolate 1.02 instead of 1.6 pixels and store 1.02 instead of 4.8 pixels per output pixel.
// CTA2 (BN = 128 tiles, F16X3): a cluster of two CTAs computes two adjacent M tiles of the same N tile with
// cta_group::2 MMAs (M = 256).  Each CTA stages its own activation tile but only HALF of the weight tile
// (rank r holds [Whi rows 64r..64r+63 ; Wlo rows 64(1-r)..]), so the weight traffic from L2 and the
// shared-memory reads of the B operand per SM are halved - the mid layers are bound by exactly that.
// Accumulator columns (per CTA, 128 lanes = its own 128 pixels): [0,64) hi*Whi[0:64] + lo*Whi[0:64],
// [64,128) hi*Wlo[64:128] + lo*Whi[64:128], [128,192) hi*Whi[64:128], [192,256) hi*Wlo[0:64].
template <int BN, int MODE, int ROW_BYTES, bool TMA_STORE, bool KHR, bool BRES = false, int UPS = 0, bool CTA2 = false>
struct TcCfg {
  static constexpr int PLANES = MODE == MODE_F16X3 ? 2 : 1;
  static constexpr int ES = MODE == MODE_F16X3 ? 2 : 4;
  static constexpr int KC_ = ROW_BYTES / ES;
  static constexpr int A_BYTES = (KHR ? 192 : 128) * ROW_BYTES;           // capacity; KHR boxes are <= 192 rows
  static constexpr int B_BYTES = (KHR ? 3 : 1) * (CTA2 ? BN / 2 : BN) * ROW_BYTES;
  // UPS == 1: rolling rows of the 2x-upsampled input, interpolated by producer warps (de_conv4_0)
  // UPS == 2: rolling rows of a full-resolution input, one TMA box per row and plane, epilogue = the two heads
  static constexpr int UPS_PX = 130;                                       // pixels per ring row (one zero column each side)
  static constexpr int UPS_PITCH = UPS == 2 ? 136 : UPS_PX;                // ring rows per slot (TMA destinations are 512-byte aligned)
  static constexpr int UPS_PLANE = UPS_PITCH * ROW_BYTES;                  // bytes of one ring slot, one plane
  static constexpr int STAGE = UPS ? UPS_PLANE * PLANES : (A_BYTES + (BRES ? 0 : B_BYTES)) * PLANES;
  static constexpr int RES = BRES ? 3 * PLANES * B_BYTES : 0;               // resident weights (all 3 kw)
  static constexpr int OUT_ROW = 32 * ES;                                  // bytes per staged row (32 columns)
  static constexpr int OUT_BUF = TMA_STORE ? PLANES * 128 * OUT_ROW : 0;   // one staging buffer
  static constexpr int BAR_BYTES = UPS ? 512 : 256;                        // mbarriers (+ the TMEM address slot)
  static constexpr int MISC = 1024 /*align*/ + BAR_BYTES + BN * 8;
  static constexpr int UPS_WARPS = UPS == 1 ? 8 : 0;                       // interpolating producer warps
  // epilogue warps: two per TMEM lane quarter for tiles of >= 2 column chunks (each takes every other 32-column
  // chunk: the accumulator of a single-wave launch drains in half the time), one per quarter otherwise
  static constexpr int EPI_SETS = (BN >= 64 && !UPS) ? 2 : 1;
  static constexpr int EPI_WARPS = 4 * EPI_SETS;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + 32 * UPS_WARPS;
  static constexpr int LR_BYTES = 0;
  static constexpr int NST_RAW = (227 * 1024 - MISC - 2 * OUT_BUF - RES - LR_BYTES) / STAGE;
  static constexpr int NST = NST_RAW > 8 ? 8 : NST_RAW;                    // UPS: ring of upsampled rows (multiple of 4:
                                                                           // keeps the lo-plane ring 512-byte aligned)
  static constexpr int KC = ROW_BYTES / ES;                                // channels per K-step
  static constexpr int MMA_PER_TILE = ROW_BYTES / 32;                      // UMMA_K spans 32 bytes
  // Every MMA re-reads its 128 x 32 B slice of A from shared memory, which is what bounds the narrow
  // tiles.  So hi*Whi and hi*Wlo are issued as ONE MMA against the stacked weight tile [Whi; Wlo]
  // (the lo tile directly follows the hi tile in shared memory; N = 2*BN, two column blocks of the
  // accumulator that the epilogue adds) and only lo*Whi needs a second MMA: A is read twice per
  // K-slice instead of three times, and 2 instead of 3 MMAs are issued.
  static constexpr bool STACK = MODE == MODE_F16X3;
  static constexpr int ACC_COLS = STACK ? 2 * BN : BN;                     // accumulator columns per buffer
  // two accumulator buffers; UPS (tap-stacked scheme): six output-row blocks in each of two regions (hi*Whi + lo*Whi |
  // hi*Wlo) = 12 * BN columns, rounded up to a power of two
  static constexpr int TMEM_COLS = UPS ? (12 * BN <= 256 ? 256 : 512) : (2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS);
  static constexpr int SMEM = NST * STAGE + 2 * OUT_BUF + RES + LR_BYTES + MISC;
  static_assert(NST >= 2 && NST <= NST_RAW, "pipeline needs at least two stages that fit in shared memory");
  static_assert(!UPS || (KHR && BRES && MODE == MODE_F16X3 && ROW_BYTES == 64 && NST == 8), "UPS is a variant of the resident-filter kh-reuse kernel");
  static_assert(UPS != 2 || (BN == 16 && !TMA_STORE), "the heads kernel computes 2 (padded to 16) output channels and stores them itself");
  static_assert(!CTA2 || (BN == 128 && MODE == MODE_F16X3 && !KHR && !BRES && !UPS), "CTA2 is a variant of the plain 128-wide F16X3 kernel");
};

// tensor maps: a[src][plane] activations, b[plane] weights, o[plane] output (TMA_STORE only)
struct TcMaps {
  CUtensorMap a[2][2];
  CUtensorMap b[2];
  CUtensorMap o[2];
};

template <int NTHREADS>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// Persistent kernel: each CTA walks tiles t = blockIdx.x, +gridDim.x, ...; the smem ring and its
// phases run continuously across tiles, and two TMEM accumulator buffers let the MMA warp start
// tile i+1 while the epilogue warps drain tile i.
template <int BN, int MODE, int ROW_BYTES, bool TMA_STORE, bool KHR, bool BRES = false, int UPS = 0, bool CTA2 = false>
__global__ void __launch_bounds__((TcCfg<BN, MODE, ROW_BYTES, TMA_STORE, KHR, BRES, UPS, CTA2>::THREADS), 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  using Cfg = TcCfg<BN, MODE, ROW_BYTES, TMA_STORE, KHR, BRES, UPS, CTA2>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  constexpr int RING = Cfg::NST * Cfg::STAGE;
  const uint32_t stage_out = base + RING;                        // 2 staging buffers (TMA_STORE)
  uint8_t* stage_out_ptr = base_ptr + RING;
  const uint32_t res_b = base + RING + 2 * Cfg::OUT_BUF;         // resident weights (BRES)
  constexpr int AFTER = RING + 2 * Cfg::OUT_BUF + Cfg::RES + Cfg::LR_BYTES;
  const uint32_t bars = base + AFTER;                            // full[NST], empty[NST], tfull[2], tempty[2], res
  const uint32_t bar_tfull = bars + 8 * (2 * Cfg::NST), bar_tempty = bar_tfull + 16, bar_res = bar_tempty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + AFTER + 8 * (2 * Cfg::NST + 5));
  float* s_scale = reinterpret_cast<float*>(base_ptr + AFTER + Cfg::BAR_BYTES);
  float* s_shift = s_scale + BN;
  // UPS, tap-stacked scheme: per output-row block one "complete" and one "drained + zeroed" barrier, and one for the
  // initial zero fill of the accumulator
  const uint32_t bar_bfull = bars + 8 * (2 * Cfg::NST + 8), bar_bempty = bar_bfull + 48, bar_zero = bar_bempty + 48;

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  // CTA2: cluster c = blockIdx.x / 2 walks pair-tiles; CTA rank r takes M tile 2*mp + r of pair-tile (mp, nt)
  const uint32_t rank = CTA2 ? uniform(cluster_ctarank()) : 0u;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // UPS: the CTA owns the contiguous output rows [ups_r0, ups_r1) (global row index = image * H + y)
  const int ups_per = UPS ? (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ups_r0 = UPS ? min((int)blockIdx.x * ups_per, p.total_tiles) : 0, ups_r1 = UPS ? min(ups_r0 + ups_per, p.total_tiles) : 0;
  // weight multicast (p.wmc): the two CTAs of a cluster run the same number of (tile, K-step) ring steps - a CTA whose
  // last tile does not exist repeats the layer's last tile (the same values are stored twice)
  const bool wmc = KHR && !BRES && !UPS && !CTA2 && p.wmc != 0;
  const uint32_t wrank = wmc ? uniform(cluster_ctarank()) : 0u;
  const int iter_first = wmc ? (tile0 & ~1) : tile0;
  const int n_iter = UPS || iter_first >= p.total_tiles ? 0 : (p.total_tiles - iter_first + tile_step - 1) / tile_step;
  const int cin = p.c0 + p.c1;
  const int cchunks = cin / Cfg::KC;
  const int ksteps = ((KHR ? 3 : p.taps) * cchunks) / p.ksplit;      // K-steps of one tile (of one split)
  const int tiles_per_group = p.tiles_x * p.tiles_y;

  const bool tl_on = (p.dbg & 256) && blockIdx.x == 0;
  if (tl_on && threadIdx.x == 0) {
    const unsigned long long slot = atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg_buf + kTimelineSlots * 8), 1ull);
    tmem_slot[1] = (uint32_t)slot;
    if (slot < kTimelineSlots) p.dbg_buf[slot * 8 + 0] = globaltimer_ns();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::NST; ++i) {
      mbar_init(bars + 8 * i, UPS == 1 ? Cfg::UPS_WARPS : 1);     // full: the TMA thread, or one arrival per producer warp
      mbar_init(bars + 8 * (Cfg::NST + i), wmc ? 2 : 1);          // (weight multicast: freed by both CTAs' MMAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, (CTA2 ? 2 : 1) * Cfg::EPI_WARPS);   // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_init(bar_res, 1);
    if (UPS) {
      for (int i = 0; i < 6; ++i) {
        mbar_init(bar_bfull + 8 * i, 1);
        mbar_init(bar_bempty + 8 * i, Cfg::EPI_WARPS);
      }
      mbar_init(bar_zero, Cfg::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (UPS != 1) tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0]);
    if (TMA_STORE) tma_prefetch_desc(&maps.o[0]);
  }
  if (warp == 1) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2 || wmc) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem = uniform(*tmem_slot);
  const uint32_t tl_slot = tl_on ? uniform(tmem_slot[1]) : 0u;
  const bool tl = tl_on && tl_slot < (uint32_t)kTimelineSlots;
  long long* const tl_buf = p.dbg_buf + (size_t)tl_slot * 8;
  if (tl && threadIdx.x == 0) tl_buf[1] = globaltimer_ns();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) may overlap the tail of the previous kernel in the stream; no global memory is
  // touched before this point.  Let the next kernel start its own prologue as early as possible.
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) may overlap the tail of the previous kernel in the stream; no global memory is
  // touched before this point.  Let the next kernel start its own prologue as early as possible.
  // (Issuing the weight loads of the first ring stages BEFORE the wait and fetching the residual tile before the
  // accumulator wait were both measured - same-box A/B, tools/_old builds - and bought nothing: the launch is not
  // waiting for those latencies)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    if (BRES) {      // the whole filter (3 kw x all kh x cin x cout, both planes) once per CTA
      if (elect_one()) {
        mbar_expect_tx(bar_res, Cfg::RES);
        for (int kw = 0; kw < 3; ++kw) {
          if (UPS && p.nstack) {
            // per kw: [Whi: kh0, kh1, kh2 rows][Wlo: kh0, kh1, kh2 rows] - the kh slices of one plane are contiguous
            tma_load_4d(res_b + kw * Cfg::PLANES * Cfg::B_BYTES, &maps.b[0], bar_res, 0, 0, 0, kw);
            tma_load_4d(res_b + kw * Cfg::PLANES * Cfg::B_BYTES + Cfg::B_BYTES, &maps.b[1], bar_res, 0, 0, 0, kw);
          } else {
            tma_load_4d(res_b + kw * Cfg::PLANES * Cfg::B_BYTES, &maps.b[0], bar_res, 0, 0, 0, kw);
          }
        }
      }
      __syncwarp();
    }
    if (UPS == 2) {
      // rolling rows: one box {32 channels, 130 pixels from x = -1, 1 row} per plane and produced row; rows -1
      // and H and the two border columns are out of bounds = zero-filled by TMA = the conv padding
      uint32_t pc = 0;
      int g = ups_r0;
      while (g < ups_r1) {
        const int img = g / p.H, ys = g - img * p.H;
        const int ye = min(p.H, ys + (ups_r1 - g));
        for (int Y = ys - 1; Y <= ye; ++Y, ++pc) {
          const uint32_t slot = pc % Cfg::NST;
          mbar_wait(bars + 8 * (Cfg::NST + slot), ((pc / Cfg::NST) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bars + 8 * slot, (uint32_t)(Cfg::PLANES * Cfg::UPS_PX * ROW_BYTES));
#pragma unroll
            for (int pl = 0; pl < Cfg::PLANES; ++pl)
              tma_load_4d(base + (uint32_t)((pl * Cfg::NST + slot) * Cfg::UPS_PLANE), &maps.a[0][pl], bars + 8 * slot, 0, -1, Y, img);
          }
          __syncwarp();
        }
        g += ye - ys;
      }
    }
    // (no divisions per K-step; ring stage / phase advanced incrementally)
    uint32_t st = 0, ph = 0;
    const int ntap = KHR ? 3 : p.taps;
    constexpr bool ld_a = true, ld_b = !BRES;
    const uint32_t a_bytes = KHR ? (uint32_t)((p.BH + 2) * p.BW * ROW_BYTES) : (uint32_t)Cfg::A_BYTES;
    const uint32_t tx = (uint32_t)Cfg::PLANES * ((ld_a ? a_bytes : 0u) + (ld_b ? (uint32_t)Cfg::B_BYTES : 0u));
    // weight operand of one K-step into ring stage `stg` (the transaction bytes of the whole stage, activations
    // included, are announced by whoever calls this first for the stage)
    auto load_b = [&](uint32_t stg, int n0, int tap, int cq) {
      const uint32_t full = bars + 8 * stg;
      const uint32_t sb = base + stg * Cfg::STAGE + Cfg::PLANES * Cfg::A_BYTES;
      if (CTA2) { if (rank == 0) mbar_expect_tx(full, 2 * tx); } else mbar_expect_tx(full, tx);
      if (KHR && wmc) {
        // the stage's weights = [kh][Whi rows ; Wlo rows] = six half boxes {kc, BN rows, one kh}: this CTA fetches
        // three of them and writes each into BOTH CTAs' stage (their bytes count on both full barriers)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int hb = (int)wrank * 3 + j, kh = hb >> 1, half = hb & 1;
          tma_load_4d_mc(sb + (uint32_t)((kh * 2 + half) * BN * ROW_BYTES), &maps.b[1], full, cq * Cfg::KC, half * BN, kh, tap, (uint16_t)3);
        }
      } else if (KHR) {
        // one box = {kc, stacked [Whi; Wlo] rows, all 3 kh, this kw}
        tma_load_4d(sb, &maps.b[0], full, cq * Cfg::KC, 0, 0, tap);
      } else {
        const int wk = tap * cin;
#pragma unroll
        for (int pl = 0; pl < Cfg::PLANES; ++pl) {
          // CTA2, rank r: [Whi rows 64r.. ; Wlo rows 64(1-r)..]
          if (CTA2) tma_load_2d_2sm(sb + pl * Cfg::B_BYTES, &maps.b[pl], full, wk + cq * Cfg::KC,
                                    n0 + (BN / 2) * (pl == 0 ? (int)rank : 1 - (int)rank));
          else tma_load_2d(sb + pl * Cfg::B_BYTES, &maps.b[pl], full, wk + cq * Cfg::KC, n0);
        }
      }
    };
    if (tl && lane == 0) tl_buf[2] = globaltimer_ns();
    for (int it = 0; it < n_iter; ++it) {
      const int t = min(tile0 + it * tile_step, p.total_tiles - 1);
      const int sp = t % p.ksplit, tq = t / p.ksplit;            // split-K slice (ksplit == 1: tq == t)
      const int nt = tq % p.tiles_n, mt = CTA2 ? 2 * (tq / p.tiles_n) + (int)rank : tq / p.tiles_n;
      const int grp = mt / tiles_per_group, trem = mt - grp * tiles_per_group;
      const int ty = trem / p.tiles_x, tx_ = trem - ty * p.tiles_x;
      const int img0 = grp * p.BNI, y0 = ty * p.BH, x0 = tx_ * p.BW, n0 = nt * BN;
      // split-K: the slice is a range of channel chunks (of every tap)
      const int cps = cchunks / p.ksplit;
      const int cq_begin = sp * cps, cq_end = cq_begin + cps;
      int kh = 0, kw = 0;
      for (int tap = 0; tap < ntap; ++tap) {
        // KHR: `tap` is kw; the box carries rows y0-1 .. y0+BH, the weights all three kh
        const int cx = KHR ? x0 + tap - 1 : x0 * p.sx + kw - p.padx;
        const int cy = KHR ? y0 - 1 : y0 * p.sy + kh - p.pady;
        for (int cq = cq_begin; cq < cq_end; ++cq) {
          int c = cq * Cfg::KC;
          int src = 0;
          if (c >= p.c0) { src = 1; c -= p.c0; }
          const uint32_t full = bars + 8 * st;
          const uint32_t sa = base + st * Cfg::STAGE;
          mbar_wait(bars + 8 * (Cfg::NST + st), ph ^ 1);
          if (elect_one()) {
            // (CTA2: both CTAs' bytes land on the leader's barrier)
            if (ld_b) load_b(st, n0, tap, cq);
            else mbar_expect_tx(full, tx);
#pragma unroll
            for (int pl = 0; pl < Cfg::PLANES; ++pl) {
              if (CTA2) tma_load_4d_2sm(sa + pl * Cfg::A_BYTES, &maps.a[src][pl], full, c, cx, cy, img0);
              else tma_load_4d(sa + pl * Cfg::A_BYTES, &maps.a[src][pl], full, c, cx, cy, img0);
            }
          }
          __syncwarp();
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
        if (++kw == p.kdiv) { kw = 0; ++kh; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    if (rank == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format
      // (bits 7/10: 0 = F16, 2 = TF32), K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24.
      const uint32_t fmt = MODE == MODE_TF32 ? 2u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | (((CTA2 ? 256u : 128u) >> 4) << 24);
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * BN) >> 3) << 17);   // N = 2*BN
      const uint32_t idesc64 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(64 >> 3) << 17);        // N = 64
      const uint64_t dconst = umma_desc<ROW_BYTES>(0);            // descriptor without the start address
      uint32_t st = 0, ph = 0, i = 0;
      if (BRES) mbar_wait(bar_res, 0);
      if (UPS && p.nstack) {
        // Tap-stacked, input-row-stationary scheme.  An input ring row Y feeds three output rows (y = Y + 1 - kh);
        // instead of one MMA per (output row, kh) with N = cout, ONE MMA per input row multiplies the row (shifted by
        // kw) with the kh slices stacked along N - [W(kh0); W(kh1); W(kh2)], N = 3 * cout - and its three column
        // blocks land in the accumulator blocks of the three output rows.  Output row j (running index) owns block
        // j % 6 of two 6-block regions (P: hi*Whi + lo*Whi, Q: hi*Wlo) at column BN * (5 - j % 6): descending, so
        // ascending kh = ascending columns; where the three blocks wrap around the ring the MMA is split in two.
        // Blocks are zero-filled by the epilogue when it drains them, every MMA accumulates.  18 MMAs of N = 3*cout
        // per input row instead of 36 of N = cout / 2*cout: the fixed ~38-clock A read of an MMA (tools/mma_probe.cu)
        // is paid half as often.
        constexpr int C = BN;
        const uint32_t idesc_n[4] = {0u, (idesc & ~(0x3Fu << 17)) | ((uint32_t)(C >> 3) << 17),
                                     (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * C) >> 3) << 17),
                                     (idesc & ~(0x3Fu << 17)) | ((uint32_t)((3 * C) >> 3) << 17)};
        mbar_wait(bar_zero, 0);
        tc_fence_after();
        uint32_t pc = 0;                 // produced-row counter: ring slot = pc % NST
        int g = ups_r0;
        int opened = 0;                  // output rows (running index) whose block has been claimed
        while (g < ups_r1) {
          const int img = g / p.H, ys = g - img * p.H;
          const int ye = min(p.H, ys + (ups_r1 - g));
          const int obase = g - ups_r0;  // running index of output row ys
          (void)img;
          for (int Y = ys - 1; Y <= ye; ++Y, ++pc) {
            const uint32_t slot = pc % Cfg::NST;
            const bool mst = (p.dbg & 2048) && blockIdx.x == 0 && lane == 0 && pc < 512;   // timing experiments
            if (mst) p.dbg_buf[pc * 8 + 0] = clock64();
            mbar_wait(bars + 8 * slot, (pc / Cfg::NST) & 1);
            tc_fence_after();
            if (mst) p.dbg_buf[pc * 8 + 1] = clock64();
            const bool valid = Y >= 0 && Y < p.H;
            // kh range whose output row y = Y + 1 - kh lies in [ys, ye)
            const int kh_lo = max(0, Y + 2 - ye), kh_hi = min(2, Y + 1 - ys);
            if (valid && kh_lo <= kh_hi) {
              const int j_top = obase + (Y + 1 - kh_lo - ys);            // highest output row touched
              while (opened <= j_top) {                                   // claim its block: drained and zeroed?
                mbar_wait(bar_bempty + 8 * (opened % 6), ((opened / 6) & 1) ^ 1);
                ++opened;
              }
              tc_fence_after();
              if (mst) p.dbg_buf[pc * 8 + 2] = clock64();
              if (elect_one()) {
                // maximal runs of kh whose blocks are adjacent: column index 5 - j % 6 grows with kh until j % 6 == 0
                int ka = kh_lo;
                while (ka <= kh_hi) {
                  const int ja = obase + (Y + 1 - ka - ys);
                  int kb = ka;
                  while (kb < kh_hi && ((ja - (kb - ka)) % 6) != 0) ++kb;
                  const int n = kb - ka + 1;
                  const uint32_t dcol = (uint32_t)(C * (5 - ja % 6));
                  const uint32_t accP = tmem + dcol, accQ = tmem + 6 * C + dcol;
#pragma unroll
                  for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t row = slot * Cfg::UPS_PITCH + kw;
                    const uint32_t a0 = base + row * ROW_BYTES;
                    const uint32_t b0 = res_b + (uint32_t)(kw * Cfg::PLANES * Cfg::B_BYTES + ka * C * ROW_BYTES);
                    const uint64_t a_hi = dconst | ((a0 >> 4) & 0x3FFF), a_lo = dconst | (((a0 + Cfg::NST * Cfg::UPS_PLANE) >> 4) & 0x3FFF);
                    const uint64_t b_hi = dconst | ((b0 >> 4) & 0x3FFF), b_lo = dconst | (((b0 + Cfg::B_BYTES) >> 4) & 0x3FFF);
#pragma unroll
                    for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk) {
                      tc_mma<MODE>(accP, a_hi + 2 * kk, b_hi + 2 * kk, idesc_n[n], 1);   // hi*Whi
                      tc_mma<MODE>(accP, a_lo + 2 * kk, b_hi + 2 * kk, idesc_n[n], 1);   // + lo*Whi
                      tc_mma<MODE>(accQ, a_hi + 2 * kk, b_lo + 2 * kk, idesc_n[n], 1);   // hi*Wlo
                    }
                  }
                  ka = kb + 1;
                }
              }
              __syncwarp();
            }
            if (elect_one()) {
              tc_commit(bars + 8 * (Cfg::NST + slot));                   // this ring row is consumed
              // output row Y - 1 has received its last contribution (kh = 2 of row Y, or Y is past the image)
              if (Y - 1 >= ys && Y - 1 < ye) tc_commit(bar_bfull + 8 * ((obase + (Y - 1 - ys)) % 6));
            }
            __syncwarp();
            if (mst) p.dbg_buf[pc * 8 + 3] = clock64();
          }
          g += ye - ys;
        }
      } else if (UPS) {
        // produced-row counter of the top halo row of the current output row; ring slot = counter % NST
        uint32_t pb = 0;
        for (int g = ups_r0; g < ups_r1; ++g, ++i) {
          const int y = g % p.H;
          const bool seg_start = g == ups_r0 || y == 0, seg_end = g == ups_r1 - 1 || y == p.H - 1;
          const uint32_t buf = i & 1;
          mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1);
          if (seg_start) {
            mbar_wait(bars + 8 * (pb % Cfg::NST), (pb / Cfg::NST) & 1);
            mbar_wait(bars + 8 * ((pb + 1) % Cfg::NST), ((pb + 1) / Cfg::NST) & 1);
          }
          mbar_wait(bars + 8 * ((pb + 2) % Cfg::NST), ((pb + 2) / Cfg::NST) & 1);
          tc_fence_after();
          const uint32_t acc = tmem + buf * Cfg::ACC_COLS;
          if (elect_one()) {
            // same accumulation order as the tile-box kernels: kw outer, kh, then the two 16-channel slices
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                const uint32_t row = ((pb + kh) % Cfg::NST) * Cfg::UPS_PITCH + kw;       // first ring row of this tap
                const uint32_t a0 = base + row * ROW_BYTES, b0 = res_b + (uint32_t)((kw * Cfg::PLANES * 3 + kh * 2) * BN * ROW_BYTES);
                const uint64_t a_hi = dconst | ((a0 >> 4) & 0x3FFF), a_lo = dconst | (((a0 + Cfg::NST * Cfg::UPS_PLANE) >> 4) & 0x3FFF);
                const uint64_t b_st = dconst | ((b0 >> 4) & 0x3FFF);
#pragma unroll
                for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk) {
                  tc_mma<MODE>(acc, a_hi + 2 * kk, b_st + 2 * kk, idesc2, (kw | kh | kk) != 0);   // hi*Whi | hi*Wlo
                  tc_mma<MODE>(acc, a_lo + 2 * kk, b_st + 2 * kk, idesc, 1);                      // + lo*Whi
                }
              }
            }
            tc_commit(bars + 8 * (Cfg::NST + pb % Cfg::NST));            // the top halo row is no longer needed
            if (seg_end) {
              tc_commit(bars + 8 * (Cfg::NST + (pb + 1) % Cfg::NST));
              tc_commit(bars + 8 * (Cfg::NST + (pb + 2) % Cfg::NST));
            }
            tc_commit(bar_tfull + 8 * buf);
          }
          __syncwarp();
          pb += seg_end ? 3 : 1;
        }
      }
      for (int it = 0; it < n_iter; ++it, ++i) {
        const int t = min(tile0 + it * tile_step, p.total_tiles - 1);
        const uint32_t buf = i & 1;
        mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem + buf * Cfg::ACC_COLS;
        // Column grouping of the three products (p.group64, the default): identical in every tile configuration
        // so that results do not depend on the tile choice (which follows the batch size): within each block of
        // 128 output channels, channels [0,64) accumulate hi*Whi + lo*Whi in one column and hi*Wlo in the other;
        // channels [64,128) accumulate hi*Wlo + lo*Whi in one column and hi*Whi in the other (that is what the
        // cta_group::2 operand split produces); the epilogue adds the two columns.
        const bool upper = !CTA2 && BN < 128 && p.group64 && ((((t / p.ksplit) % p.tiles_n) * BN) & 64) != 0;
        const uint32_t acc_lo = acc + (upper ? BN : 0);           // where lo*Whi accumulates
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(bars + 8 * st, ph);
          tc_fence_after();
          if (tl && lane == 0 && i == 0 && ks == 0) tl_buf[3] = globaltimer_ns();
          const uint32_t sa = base + st * Cfg::STAGE;
          // BRES: one K-step per kw (cin == KC), its weights sit in the resident region
          const uint32_t sb = BRES ? res_b + (uint32_t)(ks * Cfg::PLANES * Cfg::B_BYTES) : sa + Cfg::PLANES * Cfg::A_BYTES;
          if (elect_one()) {
            if (KHR) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                // tap kh reads the halo box kh image rows further down; weights: [kh][Whi rows; Wlo rows]
                const uint32_t ao = (uint32_t)(kh * p.BW * ROW_BYTES), bo = (uint32_t)(kh * 2 * BN * ROW_BYTES);
                const uint64_t a_hi = dconst | (((sa + ao) >> 4) & 0x3FFF), a_lo = dconst | (((sa + Cfg::A_BYTES + ao) >> 4) & 0x3FFF);
                const uint64_t b_st = dconst | (((sb + bo) >> 4) & 0x3FFF);
#pragma unroll
                for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk) {
                  tc_mma<MODE>(acc, a_hi + 2 * kk, b_st + 2 * kk, idesc2, (ks | kh | kk) != 0);   // hi*Whi | hi*Wlo
                  tc_mma<MODE>(acc_lo, a_lo + 2 * kk, b_st + 2 * kk, idesc, 1);                  // + lo*Whi
                }
              }
            } else {
              const uint64_t a_hi = dconst | ((sa >> 4) & 0x3FFF), b_hi = dconst | ((sb >> 4) & 0x3FFF);
              if (MODE == MODE_TF32) {
#pragma unroll
                for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk)
                  tc_mma<MODE>(acc, a_hi + 2 * kk, b_hi + 2 * kk, idesc, (ks | kk) != 0);
              } else {
                // sb holds [Whi rows][Wlo rows] back to back = the stacked operand
                const uint64_t a_lo = dconst | (((sa + Cfg::A_BYTES) >> 4) & 0x3FFF);
#pragma unroll
                for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk) {
                  tc_mma<MODE, CTA2>(acc, a_hi + 2 * kk, b_hi + 2 * kk, idesc2, (ks | kk) != 0);   // hi*Whi | hi*Wlo
                  if (!CTA2 && BN == 128 && p.group64) {
                    // + lo*Whi: channels [0,64) onto the hi*Whi columns, [64,128) onto the hi*Wlo columns
                    tc_mma<MODE>(acc, a_lo + 2 * kk, b_hi + 2 * kk, idesc64, 1);
                    tc_mma<MODE>(acc + BN + 64, a_lo + 2 * kk, b_hi + ((64 * ROW_BYTES) >> 4) + 2 * kk, idesc64, 1);
                  } else {
                    tc_mma<MODE, CTA2>(acc_lo, a_lo + 2 * kk, b_hi + 2 * kk, idesc, 1);
                  }
                }
              }
            }
            // frees this smem stage when the MMAs retire
            if (CTA2) tc_commit_2sm(bars + 8 * (Cfg::NST + st));
            else if (wmc) tc_commit_mc(bars + 8 * (Cfg::NST + st));       // both CTAs' copies of this stage hold the peer's half
            else tc_commit(bars + 8 * (Cfg::NST + st));
            if (ks == ksteps - 1) {                         // accumulator complete
              if (CTA2) tc_commit_2sm(bar_tfull + 8 * buf); else tc_commit(bar_tfull + 8 * buf);
            }
          }
          __syncwarp();
          if (++st == Cfg::NST) { st = 0; ph ^= 1; }
        }
      }
      if (tl && lane == 0) tl_buf[4] = globaltimer_ns();
    }
  } else if (warp < 2 + Cfg::EPI_WARPS) {
    // ===================== epilogue (warps 2..5 <-> TMEM lane quarters) =====================
    const int q = warp & 3;                    // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int r = q * 32 + lane;               // accumulator row = pixel within the tile
    const int xx = r % p.BW, yy = (r / p.BW) % p.BH, ni = r / (p.BW * p.BH);
    const int et = threadIdx.x - 64;           // 0 .. 32 * EPI_WARPS - 1
    const int eset = (warp - 2) >> 2;          // which of the EPI_SETS warps of this lane quarter
    // sub-box of warp q for the bulk stores = pixels 32q .. 32q+31 of the tile (box dims chosen on the host to match)
    const int sub_x = (q * 32) % p.BW, sub_y = ((q * 32) / p.BW) % p.BH, sub_n = (q * 32) / (p.BW * p.BH);
    int last_n0 = -1;
    uint32_t i = 0, chunk_ctr = 0;
    const bool nstack = UPS && p.nstack;
    if (nstack) {                               // all twelve accumulator blocks start at zero (every MMA accumulates)
      for (int c0 = 0; c0 < 12 * BN; c0 += 32) tmem_zero32(tmem + ((uint32_t)(q * 32) << 16) + c0);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_zero);
    }
    // rolling-row kernels: tile = one image row; (image, row) advance incrementally - the general decomposition
    // below costs five integer divisions per tile, which is most of a row's epilogue time there
    int ups_img = UPS ? ups_r0 / p.H : 0, ups_y = UPS ? ups_r0 - ups_img * p.H : 0;
    const int epi_cnt = UPS ? ups_r1 - ups_r0 : n_iter;
    for (int it = 0; it < epi_cnt; ++it, ++i) {
      const int t = UPS ? ups_r0 + it : min(tile0 + it * tile_step, p.total_tiles - 1);
      int sp = 0, img0, y0, x0 = 0, n0 = 0;
      if (UPS) {
        img0 = ups_img; y0 = ups_y;
        if (++ups_y == p.H) { ups_y = 0; ++ups_img; }
      } else {
        sp = t % p.ksplit;
        const int tq = t / p.ksplit;
        const int nt = tq % p.tiles_n, mt = CTA2 ? 2 * (tq / p.tiles_n) + (int)rank : tq / p.tiles_n;
        const int grp = mt / tiles_per_group, trem = mt - grp * tiles_per_group;
        const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
        img0 = grp * p.BNI; y0 = ty * p.BH; x0 = tx * p.BW; n0 = nt * BN;
      }
      if (n0 != last_n0) {
        epi_bar<32 * Cfg::EPI_WARPS>();          // nobody still reads the previous scale/shift
        for (int j = et; j < BN; j += 32 * Cfg::EPI_WARPS) {
          s_scale[j] = (p.scale ? __ldg(&p.scale[n0 + j]) : 1.f) * p.wscale;
          s_shift[j] = p.shift ? __ldg(&p.shift[n0 + j]) : 0.f;
        }
        epi_bar<32 * Cfg::EPI_WARPS>();
        last_n0 = n0;
      }
      const uint32_t buf = i & 1;
      const bool stamp = (p.dbg & 16) && blockIdx.x == 0 && et == 0 && i < 512;
      const int img = img0 + ni;
      const bool ok = img < p.n_img;
      const size_t pix = ((size_t)(ok ? img : 0) * p.H + (y0 + yy)) * p.W + (x0 + xx);
      const size_t off = pix * p.cout + n0;
      constexpr int NCH = (BN + 32 * Cfg::EPI_SETS - 1) / (32 * Cfg::EPI_SETS);     // 32-column chunks per epilogue warp
      if (stamp) p.dbg_buf[i * 8 + 0] = clock64();
      // tap-stacked scheme: output row i owns block i % 6 of the regions P and Q (see the MMA warp)
      const uint32_t blk = i % 6u;
      const uint32_t colP = nstack ? (uint32_t)(BN * (5 - (int)blk)) : buf * Cfg::ACC_COLS;
      const uint32_t colQ = nstack ? (uint32_t)(6 * BN) + colP : colP + (uint32_t)BN;
      if (nstack) mbar_wait(bar_bfull + 8 * blk, (i / 6u) & 1u);
      else mbar_wait(bar_tfull + 8 * buf, (i >> 1) & 1);
      tc_fence_after();
      if (stamp) p.dbg_buf[i * 8 + 1] = clock64();
      if (tl && et == 0 && i == 0) tl_buf[5] = globaltimer_ns();
#pragma unroll 1
      if (UPS == 2) {
        // heads: P columns [0,16) = hi*Whi + lo*Whi, Q columns [0,16) = hi*Wlo; channel 0 = pred, 1 = weight_pred
        // (spherical_model_iterative.py:371-374: relu / sigmoid / product)
        uint32_t v[32], vq[32];
        tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + colP, v);
        if (nstack) tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + colQ, vq);
        tmem_ld_wait();
        if (nstack) {                            // drained: zero the two blocks and hand them back
          tmem_zero16(tmem + ((uint32_t)(q * 32) << 16) + colP);
          tmem_zero16(tmem + ((uint32_t)(q * 32) << 16) + colQ);
          tmem_st_wait();
        } else {
          vq[0] = v[16]; vq[1] = v[17];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(nstack ? bar_bempty + 8 * blk : bar_tempty + 8 * buf);
        if ((p.dbg & 64) || !ok) continue;
        float pr = fmaxf((__uint_as_float(v[0]) + __uint_as_float(vq[0])) * p.wscale + p.heads_bp, 0.f);
        if (p.heads_conf) {
          const float cf = 1.f / (1.f + expf(-((__uint_as_float(v[1]) + __uint_as_float(vq[1])) * p.wscale + p.heads_bc)));
          pr *= cf;
          if (p.heads_il) {
            reinterpret_cast<float2*>(p.heads_pred)[pix] = make_float2(pr, cf);
            continue;
          }
          p.heads_conf[pix] = cf;
        }
        p.heads_pred[pix] = pr;
        continue;
      }
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int cb = 32 * eset + ci * 32 * Cfg::EPI_SETS;
        if (cb >= BN) break;
        uint32_t v[32];
        tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + colP + cb, v);
        if (Cfg::STACK) {                        // second column block (hi*Wlo) of the stacked accumulator
          uint32_t v2[32];
          // CTA2: see the column map above TcCfg
          tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + (CTA2 ? colP + (cb < 64 ? 192 : 64) : colQ) + cb, v2);
          tmem_ld_wait();                        // both loads in flight together
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          tmem_ld_wait();
        }
        if (cb + 32 * Cfg::EPI_SETS >= BN) {     // this warp's share of the accumulator is read: hand it back to the MMA warp
          if (nstack) {                          // (UPS == 1: BN == 32, one chunk) zero the two drained blocks first
            tmem_zero32(tmem + ((uint32_t)(q * 32) << 16) + colP);
            tmem_zero32(tmem + ((uint32_t)(q * 32) << 16) + colQ);
            tmem_st_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (nstack) mbar_arrive(bar_bempty + 8 * blk);
            else if (CTA2 && rank != 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
            else mbar_arrive(bar_tempty + 8 * buf);
          }
        }
        if (stamp) p.dbg_buf[i * 8 + 2] = clock64();
        if (p.dbg & 64) continue;
        if (p.ksplit > 1) {                      // split-K: raw partial sums, reduced (+ bias, residual, LayerNorm) by the finish kernel
          if (ok) {
            float* dst = p.partial + ((size_t)sp * p.m_total + pix) * p.cout + n0 + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st4(dst + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
          continue;
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * s_scale[cb + j] + s_shift[cb + j];
        if (p.residual && ok) {
          if (MODE == MODE_TF32) {
            const float* rs = reinterpret_cast<const float*>(p.residual) + off + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 tt = __ldg(reinterpret_cast<const float4*>(rs + j));
              f[j] += tt.x; f[j + 1] += tt.y; f[j + 2] += tt.z; f[j + 3] += tt.w;
            }
          } else {
            const __half* rhi = reinterpret_cast<const __half*>(p.residual) + off + cb;
            const __half* rlo = rhi + p.plane;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 a = __ldg(reinterpret_cast<const uint4*>(rhi + j));
              uint4 b = __ldg(reinterpret_cast<const uint4*>(rlo + j));
              const __half2* ah = reinterpret_cast<const __half2*>(&a);
              const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                float2 x = __half22float2(ah[tt]), y = __half22float2(bh[tt]);
                f[j + 2 * tt] += x.x + y.x;
                f[j + 2 * tt + 1] += x.y + y.y;
              }
            }
          }
        }
        // one (uniform) branch per chunk, not per element: with act_fn's run-time switch inlined 32 times the
        // epilogue warp spent ~1700 clocks per chunk in taken branches
        if (p.act == OFB_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == OFB_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752440f));
        }

        if (stamp) p.dbg_buf[i * 8 + 3] = clock64();
        if (TMA_STORE) {
          // stage the 128 x 32-column chunk (swizzled like the output tensor map expects) and let
          // one thread write it with a bulk tensor store; two staging buffers alternate
          // Each epilogue warp stages and stores its own 32 pixels (a sub-box of the tile): no CTA-wide barrier,
          // the four warps drain the accumulator independently.
          // two staging buffers: alternating per chunk (one warp per quarter) or one per warp set
          const uint32_t sbuf = Cfg::EPI_SETS == 2 ? (uint32_t)eset : (chunk_ctr & 1);
          if (lane == 0) {
            if (Cfg::EPI_SETS == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          __syncwarp();                                      // this warp's slice of staging buffer `sbuf` is free again
          if (stamp) p.dbg_buf[i * 8 + 4] = clock64();
          uint8_t* dst = stage_out_ptr + sbuf * Cfg::OUT_BUF;
          if (MODE == MODE_TF32) {
            // 128-byte rows, SWIZZLE_128B: 16-byte chunk index ^= row & 7
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + r * 128 + (((j >> 2) ^ (r & 7)) << 4)) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            // 64-byte rows, SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 hi4, lo4;
              __half2* hh = reinterpret_cast<__half2*>(&hi4);
              __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                __half2 h = __floats2half2_rn(f[j + 2 * tt], f[j + 2 * tt + 1]);
                float2 hf = __half22float2(h);
                hh[tt] = h;
                ll[tt] = __floats2half2_rn(f[j + 2 * tt] - hf.x, f[j + 2 * tt + 1] - hf.y);
              }
              const int sw = ((j >> 3) ^ ((r >> 1) & 3)) << 4;
              *reinterpret_cast<uint4*>(dst + r * 64 + sw) = hi4;
              *reinterpret_cast<uint4*>(dst + 128 * 64 + r * 64 + sw) = lo4;
            }
          }
          if (stamp) p.dbg_buf[i * 8 + 5] = clock64();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (stamp) p.dbg_buf[i * 8 + 6] = clock64();
          if (lane == 0 && !(p.dbg & 128)) {
            const uint32_t src = stage_out + sbuf * Cfg::OUT_BUF + (uint32_t)(q * 32 * Cfg::OUT_ROW);
#pragma unroll
            for (int pl = 0; pl < Cfg::PLANES; ++pl)
              tma_store_4d(&maps.o[pl], src + pl * 128 * Cfg::OUT_ROW, n0 + cb, x0 + sub_x, y0 + sub_y, img0 + sub_n);   // OOB images are clipped
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (stamp) p.dbg_buf[i * 8 + 7] = clock64();
          ++chunk_ctr;
        } else if (ok && !(p.dbg & 128)) {
          if (MODE == MODE_TF32) {
            float* o = reinterpret_cast<float*>(p.out) + off + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) st4(o + j, make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
          } else {
            __half* ohi = reinterpret_cast<__half*>(p.out) + off + cb;
            __half* olo = ohi + p.plane;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 hi4, lo4;
              __half2* hh = reinterpret_cast<__half2*>(&hi4);
              __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                __half2 h = __floats2half2_rn(f[j + 2 * tt], f[j + 2 * tt + 1]);
                float2 hf = __half22float2(h);
                hh[tt] = h;
                ll[tt] = __floats2half2_rn(f[j + 2 * tt] - hf.x, f[j + 2 * tt + 1] - hf.y);
              }
              *reinterpret_cast<uint4*>(ohi + j) = hi4;
              *reinterpret_cast<uint4*>(olo + j) = lo4;
            }
          }
        }
      }
    }
    if (tl && et == 0) tl_buf[6] = globaltimer_ns();
    if (TMA_STORE && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (UPS == 1) {
    // ===================== interpolating producers (warps 6..13) =====================
    // thread = (low-res column x, 8-channel chunk ch): it owns pixels X = 2x, 2x+1 of every upsampled row.
    // F.interpolate(scale 2, bilinear, align_corners=False): row Y = 2y+dy blends low-res rows y-1+dy, y+dy
    // (clamped) with weights h0, h1; same for columns.  T(r) = the horizontally blended low-res row r (2 pixels
    // x 8 channels per thread) is cached in registers for the two rows in use: consecutive upsampled rows
    // share them, so a new low-res row is fetched only every second row.
    constexpr int NT = 32 * Cfg::UPS_WARPS, C8 = Cfg::KC / 8;
    static_assert(UPS != 1 || (NT == 64 * C8), "one producer thread per (low-res column, 8-channel chunk)");
    const int pt = threadIdx.x - (64 + 32 * Cfg::EPI_WARPS);
    const int x = pt / C8, ch = pt % C8;
    const int lh = p.H >> 1, lw = p.W >> 1;
    const __half* src_hi = reinterpret_cast<const __half*>(p.ups_src);
    const __half* src_lo = src_hi + p.ups_plane;
    const int xm = max(x - 1, 0), xp = min(x + 1, lw - 1);
    const float w1a = x == 0 ? 1.f : 0.75f, w0a = 1.f - w1a;       // X = 2x   : w0a * L[x-1] + w1a * L[x]
    const float w1b = 0.25f, w0b = 0.75f;                          // X = 2x+1 : w0b * L[x]   + w1b * L[x+1]
    uint8_t* const ring_hi = base_ptr;
    uint8_t* const ring_lo = base_ptr + Cfg::NST * Cfg::UPS_PLANE;
    // zero columns X = -1 and X = 128 of every ring row, once (the slots keep their layout)
    if (pt < Cfg::NST * 2 * C8) {
      const int slot = pt / (2 * C8), side = (pt / C8) & 1, c = pt % C8;
      const uint32_t r = slot * Cfg::UPS_PITCH + (side ? Cfg::UPS_PX - 1 : 0);
      const uint32_t off = r * ROW_BYTES + ((c ^ ((r >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(ring_hi + off) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(ring_lo + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    float ta[16], tb[16];            // T(row_a), T(row_b): [pixel 0 | pixel 1][8 channels]
    int id_a = -1, id_b = -1;        // (image * lh + low-res row) held in ta / tb
    auto load_t = [&](int img, int r, float* t) {
      const size_t rowoff = ((size_t)img * lh + r) * lw;
      float l[3][8];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int xx = k == 0 ? xm : (k == 1 ? x : xp);
        const size_t idx = (rowoff + xx) * Cfg::KC + ch * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src_hi + idx));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(src_lo + idx));
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          const float2 u = __half22float2(ah[tt]), v = __half22float2(bh[tt]);
          l[k][2 * tt] = u.x + v.x; l[k][2 * tt + 1] = u.y + v.y;
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        t[c] = w0a * l[0][c] + w1a * l[1][c];
        t[8 + c] = w0b * l[1][c] + w1b * l[2][c];
      }
    };
    uint32_t pc = 0;                 // produced-row counter
    int g = ups_r0;
    while (g < ups_r1) {
      // segment = the rows of [g, ups_r1) that lie in one image: upsampled rows ys-1 .. ye are produced
      const int img = g / p.H, ys = g - img * p.H;
      const int ye = min(p.H, ys + (ups_r1 - g));
      for (int Y = ys - 1; Y <= ye; ++Y, ++pc) {
        const uint32_t slot = pc % Cfg::NST;
        {
          // The low-resolution row that the NEXT upsampled rows will need goes into L1 now (one line per plane of
          // this thread's own column; the neighbours' threads cover x-1 / x+1): the interpolating warps pace this
          // kernel (probe: 2200 clk per row against 1400 of MMA issue), and a first-touch L2 round trip in load_t
          // every second row was most of their row time.
          const int rn = min(((Y + 2) >> 1) + 1, lh - 1);
          const size_t idx = (((size_t)img * lh + rn) * lw + x) * Cfg::KC + ch * 8;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(src_hi + idx));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(src_lo + idx));
        }
        mbar_wait(bars + 8 * (Cfg::NST + slot), ((pc / Cfg::NST) & 1) ^ 1);
        uint4 hi4[2], lo4[2];
        hi4[0] = hi4[1] = lo4[0] = lo4[1] = make_uint4(0u, 0u, 0u, 0u);
        if (Y >= 0 && Y < p.H && !(p.dbg & 32)) {
          const int y = Y >> 1, dy = Y & 1;
          const int ra = img * lh + max(y - 1 + dy, 0), rb = img * lh + min(y + dy, lh - 1);
          const float h1 = dy == 0 ? (y == 0 ? 1.f : 0.75f) : 0.25f, h0 = 1.f - h1;
          // CTA-uniform branches: the row ids depend on Y only
          if (ra != id_a) {
            if (ra == id_b) {
#pragma unroll
              for (int c = 0; c < 16; ++c) ta[c] = tb[c];
            } else {
              load_t(img, ra - img * lh, ta);
            }
            id_a = ra;
          }
          if (rb != id_b) {
            if (rb == id_a) {
#pragma unroll
              for (int c = 0; c < 16; ++c) tb[c] = ta[c];
            } else {
              load_t(img, rb - img * lh, tb);
            }
            id_b = rb;
          }
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            __half2* hh = reinterpret_cast<__half2*>(&hi4[px]);
            __half2* ll = reinterpret_cast<__half2*>(&lo4[px]);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              const float f0 = h0 * ta[px * 8 + 2 * tt] + h1 * tb[px * 8 + 2 * tt];
              const float f1 = h0 * ta[px * 8 + 2 * tt + 1] + h1 * tb[px * 8 + 2 * tt + 1];
              const __half2 hv = __floats2half2_rn(f0, f1);
              const float2 hf = __half22float2(hv);
              hh[tt] = hv;
              ll[tt] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
            }
          }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const uint32_t r = slot * Cfg::UPS_PITCH + 1 + 2 * x + px;
          const uint32_t off = r * ROW_BYTES + ((ch ^ ((r >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(ring_hi + off) = hi4[px];
          *reinterpret_cast<uint4*>(ring_lo + off) = lo4[px];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * slot);
      }
      g += ye - ys;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2 || wmc) cluster_sync_all();   // the leader's MMAs read the peer's shared memory until the very end
  if (tl && threadIdx.x == 32) tl_buf[7] = globaltimer_ns();
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int tc_make_map(CUtensorMap* m, bool half, int rank, void* addr, const cuuint64_t* dims, const cuuint32_t* box,
                int row_bytes, int spatial_stride, const cuuint64_t* byte_strides, const cuuint32_t* elem_strides) {
  EncodeTiledFn fn = encode_fn();
  OFB_CHECK(fn, "conv_tc: cuTensorMapEncodeTiled is not available from the driver");
  const int es = half ? 2 : 4;
  cuuint64_t strides[4];
  cuuint64_t acc = es;
  for (int i = 0; i < rank - 1; ++i) { acc *= dims[i]; strides[i] = byte_strides ? byte_strides[i] : acc; }
  // a strided conv reads every `spatial_stride`-th pixel: TMA traversal strides on W and H
  cuuint32_t estr[4] = {1, (cuuint32_t)spatial_stride, (cuuint32_t)spatial_stride, 1};
  if (rank != 4) estr[1] = estr[2] = 1;
  if (elem_strides) for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, addr, dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OFB_CHECK(r == CUDA_SUCCESS, "conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}
static int make_map(CUtensorMap* m, bool half, int rank, void* addr, const cuuint64_t* dims, const cuuint32_t* box,
                    int row_bytes, int spatial_stride = 1, const cuuint64_t* byte_strides = nullptr,
                    const cuuint32_t* elem_strides = nullptr) {
  return tc_make_map(m, half, rank, addr, dims, box, row_bytes, spatial_stride, byte_strides, elem_strides);
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// 2x-upsample-fused variant: 32 -> 32 channels, 3x3 stride 1, split-half format, 32 x 4 pixel tiles
static bool conv_tc_ups_supported(const ofb_conv_desc* d) {
  return d->in_fmt == OFB_FMT_SPLIT16 && d->out_fmt == OFB_FMT_SPLIT16 && d->wgt_split && d->k == 3 && d->stride == 1 &&
         d->pad == 1 && d->c0 == 32 && (!d->in1 || d->c1 == 0) && d->cout == 32 && d->w == 128 && d->h % 2 == 0;
}

bool conv_tc_supported(const ofb_conv_desc* d) {
  if (d->ups2x) return conv_tc_ups_supported(d);
  if (d->in_fmt != d->out_fmt) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  if (!((d->k == 3 && d->pad == 1) || (d->k == 1 && d->pad == 0))) return false;
  if (d->stride == 2 && ((d->h | d->w) & 1)) return false;
  if (d->in_fmt == OFB_FMT_SPLIT16 && !d->wgt_split) return false;
  int c1 = d->in1 ? d->c1 : 0;
  if (d->c0 % 32 || c1 % 32 || d->cout % 32) return false;
  if (d->cout > 128 && d->cout % 128) return false;
  const int ow = d->w / d->stride, oh = d->h / d->stride;
  if (!pow2(ow) || !pow2(oh) || ow > 128) return false;
  int bw = ow, bh = oh < 128 / bw ? oh : 128 / bw;
  if (oh % bh) return false;
  return true;
}

// Launch variants are per engine handle (TcOptions, common.cuh): the engine installs its handle's options for the
// duration of a forward (TcOptScope); operator calls made directly through the C ABI use the defaults.
static thread_local const TcOptions* t_opts = nullptr;
static TcOptions k_default_opts{};      // options of operator calls made outside an engine forward
void conv_tc_default_debug(int v) { k_default_opts.dbg = v; }
void conv_tc_default_nstack(int v) { k_default_opts.nstack = v != 0; k_default_opts.nstack_ups = v >= 2; }
const TcOptions& tc_opts() { return t_opts ? *t_opts : k_default_opts; }
TcOptScope::TcOptScope(const TcOptions* o) : prev(t_opts) { t_opts = o; }
TcOptScope::~TcOptScope() { t_opts = prev; }

static long long* g_dbg_buf = nullptr;
int conv_tc_timeline_slots() { return kTimelineSlots; }
long long* conv_tc_debug_buffer() {
  if (!g_dbg_buf) {
    cudaMalloc(&g_dbg_buf, (kTimelineSlots * 8 + 8) * sizeof(long long));
    cudaMemset(g_dbg_buf, 0, (kTimelineSlots * 8 + 8) * sizeof(long long));
  }
  return g_dbg_buf;
}

// which instantiation the calling thread's last conv_tc() selected (tests assert that a shape really reaches the
// kernel they mean to cover): ofb_last_conv_variant()
static thread_local char t_variant[64] = "";
const char* conv_tc_last_variant() { return t_variant; }

// per-device state: SM count, and whether a kernel instantiation has had its dynamic-shared-memory opt-in applied
// on that device (function attributes are per device)
constexpr int kMaxDevices = 64;
static int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
static int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = cur_device();
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

template <int BN, int MODE, int ROW_BYTES, bool TMA_STORE, bool KHR, bool BRES = false, int UPS = 0, bool CTA2 = false>
static int launch_tc(const TcMaps& maps, const TcParams& p, cudaStream_t s) {
  using Cfg = TcCfg<BN, MODE, ROW_BYTES, TMA_STORE, KHR, BRES, UPS, CTA2>;
  static bool attr[kMaxDevices] = {false};
  const int dev = cur_device();
  if (!attr[dev]) {
    OFB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MODE, ROW_BYTES, TMA_STORE, KHR, BRES, UPS, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr[dev] = true;
  }
  // persistent: one CTA per SM (CTA2: one CTA pair per TPC, total_tiles counts pair-tiles)
  const int units = (CTA2 ? num_sms() / 2 : num_sms()) / tc_opts().sm_share;
  int grid = (p.total_tiles < units ? p.total_tiles : units) * (CTA2 ? 2 : 1);
  const bool wmc = KHR && !BRES && !UPS && !CTA2 && p.wmc;
  if (wmc) grid &= ~1;                        // clusters of two
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (tc_opts().pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CTA2 || wmc) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 2; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs; cfg.numAttrs = na;
  OFB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, MODE, ROW_BYTES, TMA_STORE, KHR, BRES, UPS, CTA2>, maps, p));
  OFB_LAUNCH_CHECK();
  return 0;
}

template <int MODE, int ROW_BYTES>
static int launch_bn(int bn, bool khr, bool cta2, const TcMaps& maps, const TcParams& p, cudaStream_t s) {
  if (bn == 128 && cta2) {
    if (MODE == MODE_F16X3 && ROW_BYTES == 128) {
      if (tc_opts().store128) return launch_tc<128, MODE_F16X3, 128, true, false, false, false, true>(maps, p, s);
      return launch_tc<128, MODE_F16X3, 128, false, false, false, false, true>(maps, p, s);
    }
    OFB_CHECK(false, "conv_tc: CTA pairs need split-half operands with 128-byte rows");
  }
  if (bn == 128) {
    if (tc_opts().store128) return launch_tc<128, MODE, ROW_BYTES, true, false>(maps, p, s);
    return launch_tc<128, MODE, ROW_BYTES, false, false>(maps, p, s);
  }
  if (MODE == MODE_F16X3 && khr) {
    if (p.ups_src) {
      if (ROW_BYTES == 64) {
        if (tc_opts().direct32) return launch_tc<32, MODE_F16X3, 64, false, true, true, true>(maps, p, s);
        return launch_tc<32, MODE_F16X3, 64, true, true, true, true>(maps, p, s);
      }
      OFB_CHECK(false, "conv_tc: fused upsample needs 64-byte rows");
    }
    if (bn == 64) return launch_tc<64, MODE_F16X3, ROW_BYTES, true, true>(maps, p, s);
    // 32 -> 32 channels: the whole filter stays resident in shared memory
    if (ROW_BYTES == 64 && p.c0 + p.c1 == 32 && p.cout == 32) {
      if (tc_opts().direct32) return launch_tc<32, MODE_F16X3, 64, false, true, true>(maps, p, s);
      return launch_tc<32, MODE_F16X3, 64, true, true, true>(maps, p, s);
    }
    if (tc_opts().direct32) return launch_tc<32, MODE_F16X3, ROW_BYTES, false, true>(maps, p, s);
    return launch_tc<32, MODE_F16X3, ROW_BYTES, true, true>(maps, p, s);
  }
  if (bn == 64) return launch_tc<64, MODE, ROW_BYTES, true, false>(maps, p, s);
  return launch_tc<32, MODE, ROW_BYTES, true, false>(maps, p, s);
}

// everything a launch needs: tensor maps, parameters and the tile configuration chosen for the layer
struct TcPlan { TcMaps maps; TcParams p; int bn, row_bytes; bool khr, cta2, split; };

static int tc_prepare(const ofb_conv_desc* d, TcPlan& plan) {
  OFB_CHECK(conv_tc_supported(d), "conv_tc: unsupported shape");
  const bool split = d->in_fmt == OFB_FMT_SPLIT16;
  const int c1 = d->in1 ? d->c1 : 0, cin = d->c0 + c1;
  const int es = split ? 2 : 4;
  // channels per K-step: a 128-byte row when every source allows it, else a 64-byte row
  int row_bytes = 128;
  if ((d->c0 * es) % 128 || (c1 * es) % 128) row_bytes = 64;
  // experiment (option khr_row64, off): kh-reuse layers with 64 output channels have 96 KB stages = a 2-stage
  // ring; 64-byte rows give 4 stages of half the size, but measured 3-10 % slower (TMA on 64-byte rows)
  if (tc_opts().khr_row64 && split && d->k == 3 && d->stride == 1 && d->cout == 64 && d->w >= 16 && d->h >= 8) row_bytes = 64;
  const int kc = row_bytes / es;
  OFB_CHECK(d->c0 % kc == 0 && c1 % kc == 0, "conv_tc: channel counts (%d,%d) not divisible by %d", d->c0, c1, kc);
  TcParams p{};
  const int ow = d->w / d->stride, oh = d->h / d->stride;
  p.n_img = d->n; p.H = oh; p.W = ow; p.c0 = d->c0; p.c1 = c1; p.cout = d->cout; p.k = d->k; p.pad = d->pad;
  p.stride = d->stride;
  p.taps = d->k * d->k; p.kdiv = d->k; p.sx = p.sy = d->stride; p.padx = p.pady = d->pad;
  p.BW = ow; p.BH = oh < 128 / p.BW ? oh : 128 / p.BW; p.BNI = 128 / (p.BW * p.BH);
  p.tiles_x = ow / p.BW; p.tiles_y = oh / p.BH;
  p.scale = d->scale; p.shift = d->shift; p.wscale = split ? d->wgt_unscale : 1.f;
  p.residual = d->residual; p.out = d->out; p.act = d->act;
  p.plane = (long long)d->n * oh * ow * d->cout;
  p.dbg = tc_opts().dbg;
  p.dbg_buf = (tc_opts().dbg & (16 | 256 | 2048)) ? conv_tc_debug_buffer() : nullptr;
  p.group64 = tc_opts().cta2 ? 1 : 0;
  const int S = d->ksplit > 1 ? d->ksplit : 1;
  OFB_CHECK(S == 1 || (split && (d->k == 1 || (d->k == 3 && d->cout > 64)) && d->stride == 1 && !d->ups2x && d->partial && (cin / kc) % S == 0),
            "conv_tc: split-K needs a split-half 1x1 layer (or a 3x3 one with cout > 64), a partial buffer and %d K-chunks divisible by %d", cin / kc, S);
  p.ksplit = S; p.partial = d->partial; p.m_total = (long long)d->n * oh * ow;
  const int groups = (d->n + p.BNI - 1) / p.BNI;
  // kh-reuse tiling for the narrow 3x3 layers: 32x4 / 16x8 pixel tiles inside one image.  Decided from the
  // layer shape only, never from the batch size, so results stay batch-invariant (it accumulates the taps
  // in a different order).
  const bool khr = split && d->k == 3 && d->stride == 1 && d->cout <= 64 && ow >= 16 && oh >= 8;
  // widest N tile unless that leaves most SMs without a tile
  int bn = d->cout >= 128 ? 128 : d->cout;
  // (only for really small problems such as the token linears: narrow tiles re-read the A tile more often)
  while (!khr && bn > 32 && (long long)groups * p.tiles_x * p.tiles_y * (d->cout / bn) * S < num_sms() / tc_opts().fill_div) bn >>= 1;
  int groups_k = groups;
  if (khr) {
    // 16 x 8 pixel tiles: the halo box is 16 x 10 = 1.25 x the tile (32 x 4 tiles: 32 x 6 = 1.5 x)
    p.BW = ow < tc_opts().khr_bw ? ow : tc_opts().khr_bw; p.BH = 128 / p.BW; p.BNI = 1;
    p.tiles_x = ow / p.BW; p.tiles_y = oh / p.BH;
    groups_k = d->n;
  }
  p.tiles_n = d->cout / bn;
  p.total_tiles = groups_k * p.tiles_x * p.tiles_y * p.tiles_n * S;
  // CTA pairs (cta_group::2): two adjacent M tiles share one weight tile.  Decided from the layer shape only.
  // (split-K 3x3 layers keep the pair kernel; the split-K linears keep the single-CTA tiles they were tuned with)
  const bool cta2 = tc_opts().cta2 && split && bn == 128 && !khr && row_bytes == 128 && (S == 1 || d->k == 3);
  if (cta2) p.total_tiles = ((groups_k * p.tiles_x * p.tiles_y + 1) / 2) * p.tiles_n * S;

  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int planes = split ? 2 : 1;
  if (d->ups2x) {
    // rolling-row kernel: tile = one image row of 128 pixels, a CTA walks a contiguous range of rows
    OFB_CHECK(khr && ow == 128, "conv_tc: fused upsample needs 128-pixel rows");
    p.BW = 128; p.BH = 1; p.BNI = 1; p.tiles_x = 1; p.tiles_y = oh;
    p.total_tiles = d->n * oh;
    p.ups_src = d->in0;
    p.ups_plane = (long long)d->n * (d->h / 2) * (d->w / 2) * d->c0;
  }
  for (int src = 0; src < 2 && !d->ups2x; ++src) {       // (the fused-upsample producers read in0 directly)
    const void* ptr = src == 0 ? d->in0 : d->in1;
    int c = src == 0 ? d->c0 : c1;
    if (!ptr || c == 0) { ptr = d->in0; c = d->c0; }      // unused map: keep it valid
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(p.BW * d->stride), (cuuint32_t)((khr ? p.BH + 2 : p.BH) * d->stride), (cuuint32_t)p.BNI};
    for (int pl = 0; pl < planes; ++pl) {
      char* a = (char*)ptr + (size_t)pl * d->n * d->h * d->w * c * es;
      if (make_map(&maps.a[src][pl], split, 4, a, dims, box, row_bytes, d->stride)) return -1;
    }
  }
  // The tap-stacked scheme is used by the heads only: for this 32-channel layer 18 MMAs of N = 96 cost what 36 of
  // N = 64 / 32 cost (measured with tools/probe_rolling.py: 1630 clk of issue per input row, 172 vs 162 us per launch
  // with the larger TMEM footprint and the block zero fills), for the heads' N = 48 they are cheaper (123 vs 135 us).
  p.nstack = (d->ups2x && tc_opts().nstack && tc_opts().nstack_ups) ? 1 : 0;
  // weight multicast: kh-reuse layers whose weights travel through the ring (everything but the 32 -> 32 layers)
  p.wmc = (khr && !d->ups2x && tc_opts().wmc && !(row_bytes == 64 && cin == 32 && d->cout == 32) && p.total_tiles >= 2) ? 1 : 0;
  if (khr && p.nstack) {
    // tap-stacked rolling-row kernel: per plane one map over (cout, kh, kw, cin) seen as dims {cin, cout, kh, kw};
    // one box = {kc, bn rows, all 3 kh, one kw} -> shared memory [kh][bn rows] = the kh slices of a plane stacked
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)d->cout, 3, 3};
    cuuint64_t bstr[3] = {(cuuint64_t)9 * cin * es, (cuuint64_t)3 * cin * es, (cuuint64_t)cin * es};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)bn, 3u, 1u};
    for (int pl = 0; pl < 2; ++pl)
      if (make_map(&maps.b[pl], split, 4, (char*)d->wgt_split + (size_t)pl * d->cout * 9 * cin * es, dims, box, row_bytes, 1, bstr)) return -1;
  } else if (khr) {
    // weights: the hi plane (cout, kh, kw, cin) is followed by the lo plane, i.e. one (2*cout, kh, kw, cin)
    // tensor = the stacked [Whi; Wlo] operand.  Seen as dims {cin, 2*cout, kh, kw}: one box = {kc, 2*bn rows,
    // all 3 kh, one kw} -> shared memory [kh][Whi rows; Wlo rows]
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)2 * d->cout, 3, 3};
    cuuint64_t bstr[3] = {(cuuint64_t)9 * cin * es, (cuuint64_t)3 * cin * es, (cuuint64_t)cin * es};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(2 * bn), 3u, 1u};
    if (make_map(&maps.b[0], split, 4, (char*)d->wgt_split, dims, box, row_bytes, 1, bstr)) return -1;
    cuuint32_t hbox[4] = {(cuuint32_t)kc, (cuuint32_t)bn, 1u, 1u};        // half of one kh slice (weight multicast)
    if (make_map(&maps.b[1], split, 4, (char*)d->wgt_split, dims, hbox, row_bytes, 1, bstr)) return -1;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)d->k * d->k * cin, (cuuint64_t)d->cout};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(cta2 ? bn / 2 : bn)};      // a CTA pair splits the weight tile
    for (int pl = 0; pl < planes; ++pl) {
      char* a = split ? (char*)d->wgt_split + (size_t)pl * d->cout * d->k * d->k * cin * 2 : (char*)d->wgt;
      if (make_map(&maps.b[pl], split, 2, a, dims, box, row_bytes)) return -1;
    }
  }
  if (bn < 128 || tc_opts().store128) {      // output tensor maps for the bulk-store epilogue: box = 32 columns x the pixel box
    cuuint64_t dims[4] = {(cuuint64_t)d->cout, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)d->n};
    // one store per epilogue warp: 32 consecutive pixels of the tile
    const int sbw = p.BW < 32 ? p.BW : 32, sbh = 32 / sbw < p.BH ? 32 / sbw : p.BH, sbn = 32 / (sbw * sbh);
    cuuint32_t box[4] = {32u, (cuuint32_t)sbw, (cuuint32_t)sbh, (cuuint32_t)sbn};
    for (int pl = 0; pl < planes; ++pl) {
      char* a = (char*)d->out + (size_t)pl * p.plane * es;
      if (make_map(&maps.o[pl], split, 4, a, dims, box, 32 * es)) return -1;
    }
  }
  snprintf(t_variant, sizeof(t_variant), "%s", cta2 ? "cta2" : (d->ups2x ? "ups" : (khr ? "khr" : (S > 1 ? "splitk" : "plain"))));
  plan.maps = maps; plan.p = p; plan.bn = bn; plan.row_bytes = row_bytes; plan.khr = khr; plan.cta2 = cta2; plan.split = split;
  return 0;
}

int conv_tc(const ofb_conv_desc* d, cudaStream_t s) {
  TcPlan pl;
  if (tc_prepare(d, pl)) return -1;
  if (pl.split) {
    if (pl.row_bytes == 128) return launch_bn<MODE_F16X3, 128>(pl.bn, pl.khr, pl.cta2, pl.maps, pl.p, s);
    return launch_bn<MODE_F16X3, 64>(pl.bn, pl.khr, false, pl.maps, pl.p, s);
  }
  if (pl.row_bytes == 128) return launch_bn<MODE_TF32, 128>(pl.bn, false, false, pl.maps, pl.p, s);
  return launch_bn<MODE_TF32, 64>(pl.bn, false, false, pl.maps, pl.p, s);
}

// ------------------------------------------------------------------ image-stationary layer chains
// A run of consecutive 3x3 stride-1 convs of one encoder stage (layer2: seven 128 -> 128 convs at 16x16) executed
// by ONE launch of the CTA-pair kernel's pipeline.  A pair-tile of such a layer is exactly one image (two 128-pixel
// halves) and the layer has a single N tile, so a conv depends only on what the SAME cluster wrote one layer
// earlier: every cluster walks  for layer: for its images  with the ring, the TMEM double buffer and the role warps
// running straight through.  Measured per launch of the unchained kernel at 8 panoramas (tools/timeline.py): 16.5 us
// of K loop in a 25 us period - dependency release, first-operand latency, the un-overlapped epilogue of the last
// tile, teardown and the launch gap make up the rest.  In the chain the epilogue of image A's layer l drains while
// the MMAs of image B's layer l run, and A's layer l+1 operands are already in flight when B finishes.
// Synchronisation: per local image slot one mbarrier (in both CTAs of the pair) that the epilogue warps of BOTH CTAs
// arrive on once their bulk stores of that image have completed; the producer waits for it before loading the next
// layer's activations of that image (the 3x3 halo crosses the two halves).  With several N tiles per M pair
// (layer3: 256 channels = two N tiles, handled by two clusters) the images of a pair are complete only when all
// those clusters are done: the same hand-off then goes through one global arrival counter per M pair (release /
// acquire at gpu scope, zeroed before the launch); all clusters of the launch are co-resident (<= one per SM pair)
// and a layer's tiles depend only on the previous layer, so the waits cannot form a cycle.  Residual reads use
// ld.global.cg: the buffer is rewritten during the launch, the non-coherent path could return a stale L1 line.
constexpr int CH_MAX = 12;        // layers per chain
constexpr int CH_SLOTS = 8;       // pair-tiles per cluster
struct TcLayer {
  CUtensorMap a[2], b[2], o[2];
  const float* scale; const float* shift; const void* residual;
  float wscale; int act;
};
struct TcChain {
  int L;
  int use_flags;             // dependencies cross clusters (several N tiles per M pair): global arrival counters
  unsigned int* flags;       // [M pair]: CTAs (2 per cluster x N tiles) that have finished their tile of the layers so far
  TcLayer layer[CH_MAX];
};

__global__ void __launch_bounds__((TcCfg<128, MODE_F16X3, 128, true, false, false, 0, true>::THREADS), 1)
conv_chain_kernel(const __grid_constant__ TcChain ch, const TcParams p) {
  using Cfg = TcCfg<128, MODE_F16X3, 128, true, false, false, 0, true>;
  constexpr int BN = 128, ROW_BYTES = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  constexpr int RING = Cfg::NST * Cfg::STAGE;
  const uint32_t stage_out = base + RING;
  uint8_t* stage_out_ptr = base_ptr + RING;
  constexpr int AFTER = RING + 2 * Cfg::OUT_BUF;
  const uint32_t bars = base + AFTER;                            // full[NST], empty[NST], tfull[2], tempty[2], -, dep[CH_SLOTS]
  const uint32_t bar_tfull = bars + 8 * (2 * Cfg::NST), bar_tempty = bar_tfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + AFTER + 8 * (2 * Cfg::NST + 5));
  const uint32_t bar_dep = bars + 8 * (2 * Cfg::NST + 6);
  static_assert(8 * (2 * Cfg::NST + 6 + CH_SLOTS) <= 256, "barrier block");
  float* s_scale = reinterpret_cast<float*>(base_ptr + AFTER + 256);
  float* s_shift = s_scale + BN;

  const int warp = (int)uniform(threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = uniform(cluster_ctarank());
  const int tile0 = (int)(blockIdx.x >> 1), tile_step = (int)(gridDim.x >> 1);
  const int cchunks = p.c0 / Cfg::KC;
  const int ksteps = p.taps * cchunks;
  const int tiles_per_group = p.tiles_x * p.tiles_y;
  const int L = ch.L;
  // timing experiments ("tc_debug" & 256): CTA 0 stamps %globaltimer per (layer, slot): 0 dependency released,
  // 1 first operands landed, 2 last MMA issued, 3 accumulator complete, 4 stores issued, 5 stores complete + arrive
  const bool tl = (p.dbg & 256) && blockIdx.x == 0;
  long long* const tlb = p.dbg_buf + 512 * 8;      // rows [512, 512 + CH_MAX * CH_SLOTS) of the stamp buffer

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < Cfg::NST; ++i) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (Cfg::NST + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 2 * Cfg::EPI_WARPS);
    }
    for (int i = 0; i < CH_SLOTS; ++i) mbar_init(bar_dep + 8 * i, 2);      // one arrival per CTA of the pair
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&ch.layer[0].a[0]);
    tma_prefetch_desc(&ch.layer[0].b[0]);
    tma_prefetch_desc(&ch.layer[0].o[0]);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = uniform(*tmem_slot);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t st = 0, ph = 0;
    const uint32_t tx = (uint32_t)Cfg::PLANES * ((uint32_t)Cfg::A_BYTES + (uint32_t)Cfg::B_BYTES);
    for (int l = 0; l < L; ++l) {
      const TcLayer& ly = ch.layer[l];
      int j = 0;
      for (int t = tile0; t < p.total_tiles; t += tile_step, ++j) {
        const int nt = t % p.tiles_n, mp = t / p.tiles_n, n0 = nt * BN;
        if (l > 0) {
          // the previous layer's output of these images (both halves, all channels) is complete in global memory
          if (ch.use_flags) {
            const unsigned int need = (unsigned int)(l * 2 * p.tiles_n);
            int spins = 0;
            while (true) {
              unsigned int v;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ch.flags + mp) : "memory");
              if (v >= need) break;
              if (++spins > (1 << 22)) __trap();          // a broken chain must fault, not hang the GPU
            }
          } else {
            mbar_wait(bar_dep + 8 * j, (uint32_t)((l - 1) & 1));
          }
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        if (tl && lane == 0) tlb[(l * CH_SLOTS + j) * 8 + 0] = globaltimer_ns();
        const int mt = 2 * mp + (int)rank;
        const int grp = mt / tiles_per_group, trem = mt - grp * tiles_per_group;
        const int ty = trem / p.tiles_x, tx_ = trem - ty * p.tiles_x;
        const int img0 = grp * p.BNI, y0 = ty * p.BH, x0 = tx_ * p.BW;
        int kh = 0, kw = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int cx = x0 + kw - p.padx, cy = y0 + kh - p.pady;
          const int wk = tap * p.c0;
          for (int cq = 0; cq < cchunks; ++cq) {
            const uint32_t full = bars + 8 * st;
            const uint32_t sa = base + st * Cfg::STAGE;
            const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
            mbar_wait(bars + 8 * (Cfg::NST + st), ph ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(full, 2 * tx);
#pragma unroll
              for (int pl = 0; pl < Cfg::PLANES; ++pl) {
                tma_load_4d_2sm(sa + pl * Cfg::A_BYTES, &ly.a[pl], full, cq * Cfg::KC, cx, cy, img0);
                tma_load_2d_2sm(sb + pl * Cfg::B_BYTES, &ly.b[pl], full, wk + cq * Cfg::KC,
                                n0 + (BN / 2) * (pl == 0 ? (int)rank : 1 - (int)rank));
              }
            }
            __syncwarp();
            if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          }
          if (++kw == p.kdiv) { kw = 0; ++kh; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * BN) >> 3) << 17);
      const uint64_t dconst = umma_desc<ROW_BYTES>(0);
      uint32_t st = 0, ph = 0, i = 0;
      for (int l = 0; l < L; ++l) {
        int j = 0;
        for (int t = tile0; t < p.total_tiles; t += tile_step, ++i, ++j) {
          const uint32_t buf = i & 1;
          mbar_wait(bar_tempty + 8 * buf, ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem + buf * Cfg::ACC_COLS;
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(bars + 8 * st, ph);
            tc_fence_after();
            if (tl && lane == 0 && ks == 0) tlb[(l * CH_SLOTS + j) * 8 + 1] = globaltimer_ns();
            const uint32_t sa = base + st * Cfg::STAGE;
            const uint32_t sb = sa + Cfg::PLANES * Cfg::A_BYTES;
            if (elect_one()) {
              const uint64_t a_hi = dconst | ((sa >> 4) & 0x3FFF), b_hi = dconst | ((sb >> 4) & 0x3FFF);
              const uint64_t a_lo = dconst | (((sa + Cfg::A_BYTES) >> 4) & 0x3FFF);
#pragma unroll
              for (int kk = 0; kk < Cfg::MMA_PER_TILE; ++kk) {
                tc_mma<MODE_F16X3, true>(acc, a_hi + 2 * kk, b_hi + 2 * kk, idesc2, (ks | kk) != 0);   // hi*Whi | hi*Wlo
                tc_mma<MODE_F16X3, true>(acc, a_lo + 2 * kk, b_hi + 2 * kk, idesc, 1);                 // + lo*Whi
              }
              tc_commit_2sm(bars + 8 * (Cfg::NST + st));
              if (ks == ksteps - 1) tc_commit_2sm(bar_tfull + 8 * buf);
            }
            __syncwarp();
            if (++st == Cfg::NST) { st = 0; ph ^= 1; }
          }
          if (tl && lane == 0) tlb[(l * CH_SLOTS + j) * 8 + 2] = globaltimer_ns();
        }
      }
    }
  } else {
    // ===================== epilogue (8 warps: two per TMEM lane quarter) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int xx = r % p.BW, yy = (r / p.BW) % p.BH, ni = r / (p.BW * p.BH);
    const int et = threadIdx.x - 64;
    const int eset = (warp - 2) >> 2;
    const int sub_x = (q * 32) % p.BW, sub_y = ((q * 32) / p.BW) % p.BH, sub_n = (q * 32) / (p.BW * p.BH);
    uint32_t i = 0;
    for (int l = 0; l < L; ++l) {
      const TcLayer& ly = ch.layer[l];
      int last_n0 = -1;
      int j = 0;
      for (int t = tile0; t < p.total_tiles; t += tile_step, ++i, ++j) {
        const int nt = t % p.tiles_n, mp = t / p.tiles_n, n0 = nt * BN;
        if (n0 != last_n0) {
          epi_bar<32 * Cfg::EPI_WARPS>();          // nobody still reads the previous scale / shift
          for (int jj = et; jj < BN; jj += 32 * Cfg::EPI_WARPS) {
            s_scale[jj] = (ly.scale ? __ldg(&ly.scale[n0 + jj]) : 1.f) * ly.wscale;
            s_shift[jj] = ly.shift ? __ldg(&ly.shift[n0 + jj]) : 0.f;
          }
          epi_bar<32 * Cfg::EPI_WARPS>();
          last_n0 = n0;
        }
        const int mt = 2 * mp + (int)rank;
        const int grp = mt / tiles_per_group, trem = mt - grp * tiles_per_group;
        const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
        const int img0 = grp * p.BNI, y0 = ty * p.BH, x0 = tx * p.BW;
        const uint32_t buf = i & 1;
        mbar_wait(bar_tfull + 8 * buf, (i >> 1) & 1);
        tc_fence_after();
        if (tl && et == 0) tlb[(l * CH_SLOTS + j) * 8 + 3] = globaltimer_ns();
        const int img = img0 + ni;
        const bool ok = img < p.n_img;
        const size_t pix = ((size_t)(ok ? img : 0) * p.H + (y0 + yy)) * p.W + (x0 + xx);
        const size_t off = pix * p.cout + n0;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int cb = 32 * eset + ci * 64;
          uint32_t v[32], v2[32];
          tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + buf * Cfg::ACC_COLS + cb, v);
          tmem_ld32_issue(tmem + ((uint32_t)(q * 32) << 16) + buf * Cfg::ACC_COLS + (cb < 64 ? 192 : 64) + cb, v2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(v2[k]));
          if (ci == 1) {                         // this warp's share of the accumulator is read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank != 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
              else mbar_arrive(bar_tempty + 8 * buf);
            }
          }
          float f[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]) * s_scale[cb + k] + s_shift[cb + k];
          if (ly.residual && ok) {
            const __half* rhi = reinterpret_cast<const __half*>(ly.residual) + off + cb;
            const __half* rlo = rhi + p.plane;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              const uint4 a = __ldcg(reinterpret_cast<const uint4*>(rhi + k));
              const uint4 b = __ldcg(reinterpret_cast<const uint4*>(rlo + k));
              const __half2* ah = reinterpret_cast<const __half2*>(&a);
              const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
              for (int tt = 0; tt < 4; ++tt) {
                const float2 x = __half22float2(ah[tt]), y = __half22float2(bh[tt]);
                f[k + 2 * tt] += x.x + y.x;
                f[k + 2 * tt + 1] += x.y + y.y;
              }
            }
          }
          if (ly.act == OFB_ACT_RELU) {
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = fmaxf(f[k], 0.f);
          }
          // stage this warp's 32 pixels x 32 columns (64-byte rows, SWIZZLE_64B) and store them with one bulk tensor
          // store per plane; one staging buffer per warp set
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          uint8_t* dst = stage_out_ptr + eset * Cfg::OUT_BUF;
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            uint4 hi4, lo4;
            __half2* hh = reinterpret_cast<__half2*>(&hi4);
            __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              const __half2 h = __floats2half2_rn(f[k + 2 * tt], f[k + 2 * tt + 1]);
              const float2 hf = __half22float2(h);
              hh[tt] = h;
              ll[tt] = __floats2half2_rn(f[k + 2 * tt] - hf.x, f[k + 2 * tt + 1] - hf.y);
            }
            const int sw = ((k >> 3) ^ ((r >> 1) & 3)) << 4;
            *reinterpret_cast<uint4*>(dst + r * 64 + sw) = hi4;
            *reinterpret_cast<uint4*>(dst + 128 * 64 + r * 64 + sw) = lo4;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            const uint32_t src = stage_out + eset * Cfg::OUT_BUF + (uint32_t)(q * 32 * Cfg::OUT_ROW);
#pragma unroll
            for (int pl = 0; pl < Cfg::PLANES; ++pl)
              tma_store_4d(&ly.o[pl], src + pl * 128 * Cfg::OUT_ROW, n0 + cb, x0 + sub_x, y0 + sub_y, img0 + sub_n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        // this image's rows of layer l are in global memory once every epilogue warp's bulk stores have completed:
        // release the next layer's loads of the image in both CTAs of the pair
        if (tl && et == 0) tlb[(l * CH_SLOTS + j) * 8 + 4] = globaltimer_ns();
        if (l + 1 < L) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          __syncwarp();
          epi_bar<32 * Cfg::EPI_WARPS>();
          if (et == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            if (ch.use_flags) {
              asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ch.flags + mp) : "memory");
            } else {
              mbar_arrive_remote(bar_dep + 8 * j, 0);
              mbar_arrive_remote(bar_dep + 8 * j, 1);
            }
            if (tl) tlb[(l * CH_SLOTS + j) * 8 + 5] = globaltimer_ns();
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// descs: L consecutive single-source 3x3 stride-1 convs over the same (n, h, w) with 128 -> 128 channels, each reading
// the previous one's output.  Returns 1 (nothing launched) when the chain conditions do not hold for these shapes.
int conv_tc_chain(const ofb_conv_desc* descs, int L, unsigned int* flags, int flags_capacity, cudaStream_t s) {
  if (L < 2 || L > CH_MAX || !tc_opts().cta2) return 1;
  using Cfg = TcCfg<128, MODE_F16X3, 128, true, false, false, 0, true>;
  TcChain ch;
  memset(&ch, 0, sizeof(ch));
  ch.L = L;
  TcParams p0{};
  for (int l = 0; l < L; ++l) {
    const ofb_conv_desc& d = descs[l];
    if (d.in1 || d.k != 3 || d.stride != 1 || d.ups2x || d.ksplit > 1 || !conv_tc_supported(&d)) return 1;
    if (l > 0 && (d.in0 != descs[l - 1].out || d.n != descs[0].n || d.h != descs[0].h || d.w != descs[0].w ||
                  d.c0 != descs[0].c0 || d.cout != descs[0].cout)) return 1;
    TcPlan pl;
    if (tc_prepare(&d, pl)) return -1;
    // CTA pairs whose pair-tile is a whole number of images; several N tiles need the global counters
    if (!pl.cta2 || !pl.split || pl.bn != 128 || pl.row_bytes != 128 || !tc_opts().store128) return 1;
    if (pl.p.tiles_n > 1 && (!flags || (pl.p.total_tiles / pl.p.tiles_n) > flags_capacity)) return 1;
    if ((2 * 128) % (d.h * d.w) != 0 && (d.h * d.w) % (2 * 128) != 0) return 1;
    if (d.h * d.w > 2 * 128) return 1;                   // an image larger than a pair-tile would need neighbours' rows
    if (l == 0) p0 = pl.p;
    TcLayer& ly = ch.layer[l];
    ly.a[0] = pl.maps.a[0][0]; ly.a[1] = pl.maps.a[0][1];
    ly.b[0] = pl.maps.b[0]; ly.b[1] = pl.maps.b[1];
    ly.o[0] = pl.maps.o[0]; ly.o[1] = pl.maps.o[1];
    ly.scale = d.scale; ly.shift = d.shift; ly.residual = d.residual; ly.wscale = pl.p.wscale; ly.act = d.act;
    if (d.act != OFB_ACT_RELU && d.act != OFB_ACT_NONE) return 1;
  }
  static bool attr[kMaxDevices] = {false};
  const int dev = cur_device();
  if (!attr[dev]) {
    OFB_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr[dev] = true;
  }
  const int units = (num_sms() / 2) / tc_opts().sm_share;
  const int clusters = p0.total_tiles < units ? p0.total_tiles : units;
  // more than 4 pair-tiles per cluster: the launch already amortises its overhead (measured at 32 panoramas per step,
  // 8 per cluster: 2104-2125 chained vs 2127-2131 panoramas/s unchained), separate launches are used
  if ((p0.total_tiles + clusters - 1) / clusters > 4) return 1;
  ch.use_flags = p0.tiles_n > 1 ? 1 : 0;
  ch.flags = flags;
  if (ch.use_flags) OFB_CUDA(cudaMemsetAsync(flags, 0, (size_t)(p0.total_tiles / p0.tiles_n) * sizeof(unsigned int), s));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * 2); cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = s;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (tc_opts().pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attrs[na].id = cudaLaunchAttributeClusterDimension;
  attrs[na].val.clusterDim.x = 2; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attrs; cfg.numAttrs = na;
  OFB_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, ch, p0));
  OFB_LAUNCH_CHECK();
  snprintf(t_variant, sizeof(t_variant), "chain%d", L);
  return 0;
}

// ------------------------------------------------------------------ heads on tensor cores
// pred / weight_pred (3x3, 32 -> 1 each) as ONE 3x3 conv with 16 output channels (0 = pred, 1 = weight_pred,
// 2..15 zero) on the rolling-row kernel: a CTA walks a contiguous range of 128-pixel image rows, the TMA
// producer adds one input row per output row to a ring in shared memory (1.02 instead of 1.27 x the input
// through L2, no index arithmetic), all nine taps read it through row-shifted descriptors, and the epilogue
// applies relu / sigmoid / product and writes the two float32 patch maps.  x: split-half (n, h, 128, 32).
int conv_tc_heads(const void* x, int n, int h, int w, const void* wgt_split, float wgt_unscale, float b_pred,
                  float b_conf, int confidence, float* pred_out, float* conf_out, int interleaved, cudaStream_t s) {
  OFB_CHECK(x && wgt_split && pred_out && (!confidence || conf_out), "heads_tc: null pointer");
  OFB_CHECK(!interleaved || (confidence && (reinterpret_cast<uintptr_t>(pred_out) & 7) == 0),
            "heads_tc: the interleaved (pred*conf, conf) output needs confidence != 0 and an 8-byte aligned buffer");
  OFB_CHECK(w == 128 && h >= 1, "heads_tc: rows must be 128 pixels wide (got %d)", w);
  TcParams p{};
  p.n_img = n; p.H = h; p.W = w; p.c0 = 32; p.c1 = 0; p.cout = 16; p.k = 3; p.pad = 1; p.stride = 1;
  p.taps = 9; p.kdiv = 3; p.sx = p.sy = 1; p.padx = p.pady = 1;
  p.BW = 128; p.BH = 1; p.BNI = 1; p.tiles_x = 1; p.tiles_y = h;
  p.wscale = wgt_unscale; p.act = OFB_ACT_NONE;
  p.tiles_n = 1; p.total_tiles = n * h;
  p.ksplit = 1;
  p.heads_pred = pred_out; p.heads_conf = confidence ? conf_out : nullptr; p.heads_bp = b_pred; p.heads_bc = b_conf;
  p.heads_il = interleaved ? 1 : 0;
  p.dbg = tc_opts().dbg;
  p.dbg_buf = (tc_opts().dbg & (16 | 256 | 2048)) ? conv_tc_debug_buffer() : nullptr;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const size_t plane = (size_t)n * h * w * 32;               // halves per plane
  for (int pl = 0; pl < 2; ++pl) {
    cuuint64_t dims[4] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint32_t box[4] = {32u, 130u, 1u, 1u};
    if (make_map(&maps.a[0][pl], true, 4, (char*)x + pl * plane * 2, dims, box, 64)) return -1;
    maps.a[1][pl] = maps.a[0][pl];
  }
  p.nstack = tc_opts().nstack ? 1 : 0;
  if (p.nstack) {  // per plane (16, kh, kw, 32): one box = {32 channels, 16 rows, all 3 kh, one kw} -> [kh][16 rows]
    cuuint64_t dims[4] = {32, 16, 3, 3};
    cuuint64_t bstr[3] = {(cuuint64_t)9 * 32 * 2, (cuuint64_t)3 * 32 * 2, (cuuint64_t)32 * 2};
    cuuint32_t box[4] = {32u, 16u, 3u, 1u};
    for (int pl = 0; pl < 2; ++pl)
      if (make_map(&maps.b[pl], true, 4, (char*)wgt_split + (size_t)pl * 16 * 9 * 32 * 2, dims, box, 64, 1, bstr)) return -1;
  } else {  // weights: (2*16, kh, kw, 32) = stacked [Whi; Wlo]; one box = {32 channels, 32 rows, all 3 kh, one kw}
    cuuint64_t dims[4] = {32, 32, 3, 3};
    cuuint64_t bstr[3] = {(cuuint64_t)9 * 32 * 2, (cuuint64_t)3 * 32 * 2, (cuuint64_t)32 * 2};
    cuuint32_t box[4] = {32u, 32u, 3u, 1u};
    if (make_map(&maps.b[0], true, 4, (char*)wgt_split, dims, box, 64, 1, bstr)) return -1;
    maps.b[1] = maps.b[0];
  }
  return launch_tc<16, MODE_F16X3, 64, false, true, true, 2>(maps, p, s);
}

// ------------------------------------------------------------------ stem on tensor cores
// Conv 7x7 s2 p3 3->64 as an implicit GEMM with K = 7 (kh) x 32: the input patches are stored
// split-half, 4 channels per pixel, each row padded by 4 zero pixels on both sides.  For output
// pixel (oy,ox) and tap row kh the 8 input pixels 2ox-4 .. 2ox+3 of row 2oy-3+kh are 32
// contiguous halves, so the A operand is a tensor map whose innermost dimension is that 64-byte
// window and whose second dimension (the window index ox) has a 16-byte stride: overlapping
// windows read straight out of the image by TMA, no im2col buffer.  kw = -4 carries a zero weight.
int stem_tc(const void* patches, int n, int h, int w, const void* wgt_split, float wgt_unscale, const float* scale,
            const float* shift, void* out, cudaStream_t s) {
  OFB_CHECK(patches && wgt_split && scale && shift && out, "stem_tc: null pointer");
  OFB_CHECK(h % 4 == 0 && w % 2 == 0 && w / 2 <= 128 && pow2(w / 2) && pow2(h / 2) && 128 % (w / 2) == 0,
            "stem_tc: unsupported patch size %dx%d", h, w);
  const int ow = w / 2, oh = h / 2, pitch = w + 8;           // pixels per padded row
  TcParams p{};
  p.n_img = n; p.H = oh; p.W = ow; p.c0 = 32; p.c1 = 0; p.cout = 64; p.k = 7; p.pad = 3; p.stride = 2;
  p.taps = 7; p.kdiv = 1; p.sx = 1; p.sy = 2; p.padx = 0; p.pady = 3;
  p.BW = ow; p.BH = 128 / ow; p.BNI = 1;
  OFB_CHECK(oh % p.BH == 0, "stem_tc: unsupported patch size %dx%d", h, w);
  p.tiles_x = 1; p.tiles_y = oh / p.BH;
  p.scale = scale; p.shift = shift; p.wscale = wgt_unscale; p.residual = nullptr; p.out = out; p.act = OFB_ACT_RELU;
  p.plane = (long long)n * oh * ow * 64;
  p.tiles_n = 1; p.total_tiles = n * p.tiles_y;
  p.ksplit = 1;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const size_t in_plane = (size_t)n * h * pitch * 4;         // halves per plane
  for (int pl = 0; pl < 2; ++pl) {
    cuuint64_t dims[4] = {32, (cuuint64_t)ow, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t bstr[3] = {16, (cuuint64_t)pitch * 8, (cuuint64_t)h * pitch * 8};
    cuuint32_t box[4] = {32u, (cuuint32_t)ow, (cuuint32_t)(2 * p.BH), 1u};
    cuuint32_t estr[4] = {1, 1, 2, 1};
    char* a = (char*)patches + pl * in_plane * 2;
    if (make_map(&maps.a[0][pl], true, 4, a, dims, box, 64, 1, bstr, estr)) return -1;
    maps.a[1][pl] = maps.a[0][pl];
    cuuint64_t bd[2] = {7 * 32, 64};
    cuuint32_t bb[2] = {32u, 64u};
    if (make_map(&maps.b[pl], true, 2, (char*)wgt_split + (size_t)pl * 64 * 7 * 32 * 2, bd, bb, 64)) return -1;
    cuuint64_t od[4] = {64, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)n};
    cuuint32_t ob[4] = {32u, (cuuint32_t)(p.BW < 32 ? p.BW : 32), (cuuint32_t)(p.BW < 32 ? 32 / p.BW : 1), 1u};   // per-warp sub-box
    if (make_map(&maps.o[pl], true, 4, (char*)out + (size_t)pl * p.plane * 2, od, ob, 64)) return -1;
  }
  return launch_tc<64, MODE_F16X3, 64, true, false>(maps, p, s);
}

// ------------------------------------------------------------------ attention on tensor cores
// Attention core of a Transformer_Block (model/blocks.py:50-62) on tcgen05: softmax(q k^T / sqrt(d)) v per panorama
// and head, d = 128, N <= 64 tokens per panorama.  One CTA = one head x one ROW TILE of G whole panoramas, each in
// a SLOT of `slot` = N rounded up to 16 rows (G = 128 / slot: 4 panoramas at N = 18 or 26, 2 at 46): the tile's
// 128 rows are the M dimension of both GEMMs,
//   S = Q K^T   (M = 128 query rows, N = 128 key rows, K = 128 dims)
//   O = P V     (M = 128 query rows, N = 128 dims,     K = 128 keys)
// i.e. all pairs of tokens of the tile are scored at once and the softmax keeps only the block of a row's own
// panorama (block-diagonal mask) - that wastes most of S, but one M = 128 tcgen05 tile is the smallest unit the
// tensor pipe has and the whole problem is 24 + 24 MMAs.  Slots start at multiples of 16 keys = the K extent of one
// MMA, so a panorama's keys meet the same K16 blocks wherever it sits in a tile and the other blocks contribute
// exact zeros: results do not depend on the batch composition, bit for bit.  Operands are the split-half planes of
// the fused qkv activation: Q and K slots arrive by TMA (2-D boxes of 64 dims x slot rows, 128B swizzle = the
// K-major layout tcgen05 reads), V is transposed into the same layout by the CTA's threads (V^T: rows = dims,
// K = keys), and the three products hi*hi + hi*lo + lo*hi accumulate in one TMEM accumulator (fp32-level result
// like the conv engine).  The softmax runs with one thread per query row straight out of TMEM (tcgen05.ld), writes
// P as split-half planes over the Q tiles (Q is dead once S is complete), and the O rows go back to global memory
// as split-half planes.
struct AttMaps { CUtensorMap qk[2]; };      // hi / lo plane of the (rows, 3*dim) qkv activation

constexpr int ATT_TILE = 128 * 128 * 2;     // one 128-row x 128-column fp16 operand = two 16 KB K-chunks
constexpr int ATT_SMEM = 6 * ATT_TILE + 1024 + 64 + 2 * 2 * 128 * 4;

__global__ void __launch_bounds__(256, 1)
attention_tc_kernel(const __grid_constant__ AttMaps maps, const __half* __restrict__ qkv, long long plane, int rows,
                    int N, int G, int slot, int heads, float scale, __half* __restrict__ out, long long out_plane) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  // [0] Q hi | [1] Q lo (later P hi | P lo) | [2] K hi | [3] K lo | [4] V^T hi | [5] V^T lo
  const uint32_t bar_ld = base + 6 * ATT_TILE, bar_mma = bar_ld + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + 6 * ATT_TILE + 16);
  float* red_max = reinterpret_cast<float*>(base_ptr + 6 * ATT_TILE + 64);       // [2 column halves][128 rows]
  float* red_sum = red_max + 256;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int head = blockIdx.x % heads, tile = blockIdx.x / heads;
  const int dim = heads * 128;
  const int row0 = tile * G * N;                     // first token row of the tile's first panorama
  const int pans = min(G, (rows - row0) / N);        // panoramas of this tile that exist

  if (tid == 0) {
    mbar_init(bar_ld, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&maps.qk[0]);
    tma_prefetch_desc(&maps.qk[1]);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (tid == 0) {
    // Q and K slots of this head: per plane and panorama two 64-dim chunks each.  A box is `slot` rows: the rows
    // behind a panorama's N own ones belong to the next panorama (masked below); rows beyond the activation are
    // zero-filled by TMA, so every row of the tile is initialised.
    mbar_expect_tx(bar_ld, (uint32_t)(8 * G * slot * 128));
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint32_t dst = (uint32_t)(kc * 16384 + g * slot * 128);
          tma_load_2d(base + pl * ATT_TILE + dst, &maps.qk[pl], bar_ld, head * 128 + kc * 64, row0 + g * N);
          tma_load_2d(base + (2 + pl) * ATT_TILE + dst, &maps.qk[pl], bar_ld, dim + head * 128 + kc * 64, row0 + g * N);
        }
  }
  if (G * slot < 128) {        // rows behind the last slot (N = 46: 96 of 128): never loaded, must not hold NaN patterns
    for (int i = G * slot * 128 + tid * 16; i < 128 * 128; i += 256 * 16) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        *reinterpret_cast<uint4*>(base_ptr + t * ATT_TILE + i) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(base_ptr + t * ATT_TILE + 16384 + i) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  // V^T: element (dim d, key j) of plane pl -> chunk j / 64, row d, column j % 64 (128B swizzle: 16-byte chunk ^= row & 7).
  // Thread = key column j; per step it carries 8 dims of its key.  The 32 lanes of a warp then write 32 consecutive
  // 2-byte columns of the SAME row (conflict-free), and the global loads of four steps are in flight together.
  {
    const __half* vsrc = qkv + 2 * dim + head * 128;
    const int j = tid & 127;
    const int g = j / slot, t = j - g * slot;
    const bool have = g < pans && t < N;
    const size_t o = have ? (size_t)(row0 + g * N + t) * (3 * dim) : 0;
    const uint32_t kc = (uint32_t)j >> 6, col = (uint32_t)(j & 63) * 2;          // byte column inside the 128-byte row
#pragma unroll 1
    for (int dg0 = (tid >> 7) * 8; dg0 < (tid >> 7) * 8 + 8; dg0 += 4) {
      uint4 hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        hi[q] = lo[q] = make_uint4(0u, 0u, 0u, 0u);
        if (have) {
          hi[q] = __ldg(reinterpret_cast<const uint4*>(vsrc + o + (dg0 + q) * 8));
          lo[q] = __ldg(reinterpret_cast<const uint4*>(vsrc + plane + o + (dg0 + q) * 8));
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __half* hh = reinterpret_cast<const __half*>(&hi[q]);
        const __half* ll = reinterpret_cast<const __half*>(&lo[q]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t r = (uint32_t)(dg0 + q) * 8 + u;
          const uint32_t off = kc * 16384 + r * 128 + ((((col >> 4) ^ (r & 7)) << 4) | (col & 15));
          *reinterpret_cast<__half*>(base_ptr + 4 * ATT_TILE + off) = hh[u];
          *reinterpret_cast<__half*>(base_ptr + 5 * ATT_TILE + off) = ll[u];
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // instruction descriptor: F32 accumulate, F16 operands, K-major A and B, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t dconst = umma_desc<128>(0);
  auto gemm3 = [&](uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t acc) {
    // hi*hi + hi*lo + lo*hi over K = 128 (two 64-wide chunks x four K16 slices)
#pragma unroll
    for (int kc = 0; kc < 2; ++kc) {
      const uint64_t ah = dconst | (((a_hi + kc * 16384) >> 4) & 0x3FFF), al = dconst | (((a_lo + kc * 16384) >> 4) & 0x3FFF);
      const uint64_t bh = dconst | (((b_hi + kc * 16384) >> 4) & 0x3FFF), bl = dconst | (((b_lo + kc * 16384) >> 4) & 0x3FFF);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        tc_mma<MODE_F16X3>(acc, ah + 2 * kk, bh + 2 * kk, idesc, (kc | kk) != 0);
        tc_mma<MODE_F16X3>(acc, ah + 2 * kk, bl + 2 * kk, idesc, 1);
        tc_mma<MODE_F16X3>(acc, al + 2 * kk, bh + 2 * kk, idesc, 1);
      }
    }
  };
  if (tid == 0) {
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    gemm3(base, base + ATT_TILE, base + 2 * ATT_TILE, base + 3 * ATT_TILE, tmem);          // S = Q K^T
    tc_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // softmax: two threads per query row r (TMEM lane r; warps w and w + 4 share a lane quarter), each owning 64 of
  // the 128 key columns; keys of the row's own panorama are columns [c0, c1).  Sweep 1: row maximum.  Sweep 2:
  // e = exp(s - max) written UNNORMALISED as P (values in (0, 1], the largest exactly 1: ideal for the split-half
  // planes) plus the row sum; 1 / sum is applied to the O row in the epilogue.
  const int r = tid & 127, half = tid >> 7;
  const int g = r / slot;
  const int c0 = g * slot, c1 = c0 + N;
  const bool live = g < pans && r - c0 < N;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  {
    float m = -INFINITY;
#pragma unroll 1
    for (int cb = half * 64; cb < half * 64 + 64; cb += 32) {
      uint32_t v[32];
      tmem_ld32_issue(trow + cb, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cb + j;
        if (c >= c0 && c < c1) m = fmaxf(m, __uint_as_float(v[j]) * scale);
      }
    }
    red_max[half * 128 + r] = m;
    __syncthreads();
    m = fmaxf(red_max[r], red_max[128 + r]);
    float sum = 0.f;
    uint8_t* p_hi = base_ptr, * p_lo = base_ptr + ATT_TILE;        // P over the Q tiles: K-major, two 64-key chunks
#pragma unroll 1
    for (int cb = half * 64; cb < half * 64 + 64; cb += 32) {
      uint32_t v[32];
      tmem_ld32_issue(trow + cb, v);
      tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cb + j;
        pv[j] = (live && c >= c0 && c < c1) ? expf(__uint_as_float(v[j]) * scale - m) : 0.f;
        sum += pv[j];
      }
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) {       // 16-byte chunks of 8 keys
        uint4 hi4, lo4;
        __half2* hh = reinterpret_cast<__half2*>(&hi4);
        __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float f0 = pv[q8 * 8 + 2 * u], f1 = pv[q8 * 8 + 2 * u + 1];
          const __half2 h = __floats2half2_rn(f0, f1);
          const float2 hf = __half22float2(h);
          hh[u] = h;
          ll[u] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
        }
        const uint32_t ch = (uint32_t)(cb >> 3) + q8;               // 16-byte chunk index 0..15 along the 128 keys
        const uint32_t kc = ch >> 3, c16 = ch & 7;
        const uint32_t off = kc * 16384 + (uint32_t)r * 128 + ((c16 ^ ((uint32_t)r & 7)) << 4);
        *reinterpret_cast<uint4*>(p_hi + off) = hi4;
        *reinterpret_cast<uint4*>(p_lo + off) = lo4;
      }
    }
    red_sum[half * 128 + r] = sum;
  }
  tc_fence_before();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    gemm3(base, base + ATT_TILE, base + 4 * ATT_TILE, base + 5 * ATT_TILE, tmem + 128);  // O = P V
    tc_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  // O row r (this thread's 64 dims), normalised -> out[(row0 + ..), head*128 ..] as split-half planes
  const float inv = 1.f / (red_sum[r] + red_sum[128 + r]);
#pragma unroll 1
  for (int cb = half * 64; cb < half * 64 + 64; cb += 32) {
    uint32_t v[32];
    tmem_ld32_issue(trow + 128 + cb, v);
    tmem_ld_wait();
    if (live) {
      __half* ohi = out + (size_t)(row0 + g * N + (r - c0)) * dim + head * 128 + cb;
      __half* olo = ohi + out_plane;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 hi4, lo4;
        __half2* hh = reinterpret_cast<__half2*>(&hi4);
        __half2* ll = reinterpret_cast<__half2*>(&lo4);
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          const float f0 = __uint_as_float(v[j + 2 * tt]) * inv, f1 = __uint_as_float(v[j + 2 * tt + 1]) * inv;
          const __half2 h = __floats2half2_rn(f0, f1);
          const float2 hf = __half22float2(h);
          hh[tt] = h;
          ll[tt] = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
        }
        *reinterpret_cast<uint4*>(ohi + j) = hi4;
        *reinterpret_cast<uint4*>(olo + j) = lo4;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// qkv: split-half planes of (rows, 3*heads*128) = [q | k | v] per row, rows = B*N; out: split-half (rows, heads*128)
int attention_tc(const void* qkv, int B, int N, int heads, void* out, cudaStream_t s) {
  OFB_CHECK(qkv && out && B > 0 && heads > 0, "attention_tc: bad arguments");
  OFB_CHECK(N >= 1 && N <= 64, "attention_tc: 1..64 tokens per panorama (got %d)", N);
  const int rows = B * N, dim = heads * 128, slot = (N + 15) & ~15, G = 128 / slot, tiles = (B + G - 1) / G;
  AttMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int pl = 0; pl < 2; ++pl) {
    cuuint64_t dims[2] = {(cuuint64_t)3 * dim, (cuuint64_t)rows};
    cuuint32_t box[2] = {64u, (cuuint32_t)slot};
    char* a = (char*)qkv + (size_t)pl * rows * 3 * dim * 2;
    if (make_map(&maps.qk[pl], true, 2, a, dims, box, 128)) return -1;
  }
  static bool attr[kMaxDevices] = {false};
  const int dev = cur_device();
  if (!attr[dev]) {
    OFB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles * heads); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = ATT_SMEM; cfg.stream = s;
  cudaLaunchAttribute at[1];
  int na = 0;
  if (tc_opts().pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  const __half* q = reinterpret_cast<const __half*>(qkv);
  OFB_CUDA(cudaLaunchKernelEx(&cfg, attention_tc_kernel, maps, q, (long long)rows * 3 * dim, rows, N, G, slot, heads,
                              1.f / sqrtf(128.f), reinterpret_cast<__half*>(out), (long long)rows * dim));
  OFB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------ split-half conversion
__global__ void split_kernel(const float* __restrict__ src, size_t n, float mul, __half* __restrict__ hi,
                             __half* __restrict__ lo) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src[i] * mul;
  __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}
__global__ void merge_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, size_t n,
                             float* __restrict__ dst) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = __half2float(hi[i]) + __half2float(lo[i]);
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_split_f16(const float* src, size_t n, float mul, void* dst_planes, void* stream) {
  OFB_CHECK(src && dst_planes && n > 0, "split_f16: bad arguments");
  __half* hi = reinterpret_cast<__half*>(dst_planes);
  split_kernel<<<cdiv((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(src, n, mul, hi, hi + n);
  OFB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ofb_merge_f16(const void* src_planes, size_t n, float* dst, void* stream) {
  OFB_CHECK(src_planes && dst && n > 0, "merge_f16: bad arguments");
  const __half* hi = reinterpret_cast<const __half*>(src_planes);
  merge_kernel<<<cdiv((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(hi, hi + n, n, dst);
  OFB_LAUNCH_CHECK();
  return 0;
}
