// Per-point MLP layers on the 5th-gen tensor cores (M1-M3 of SURVEY §2.2).
//
// Replaces nn.Linear / cuBLAS sgemm on the point path (pointnet.py:36-40,72-82, resnet.py:46-54,
// alto.py:63-69,123-128,164-170,248-253, pixel.py:48-58) and its autograd backward.
//
//   y[r, n] = sum_k act(x[r, k]) * w[n, k] + bias[n]      (* mask[r, n] > 0)  (+ residual[r, n])
//
// fp32 in / fp32 out with fp32-grade accuracy on the tensor cores: every operand is split into hi + lo and
// three tcgen05.mma products (lo*hi + hi*lo + hi*hi) are accumulated in fp32 in tensor memory -- as TF32
// terms ("3xTF32", kind::tf32) or, for the wide layers, as fp16 terms of power-of-two scaled operands
// ("3xFP16", kind::f16, twice the rate; see f16_scale_exp).  Kernels in this file:
//   linear_x3_persistent_kernel   forward / input-gradient GEMM and 3x3 convolution (implicit GEMM), the
//                                 product path: one persistent CTA per SM, TMA producer warp, MMA warp,
//                                 operand-split warps (x -> TMEM), two epilogue groups, double-buffered
//                                 TMEM accumulators
//   linear_tf32x3_kernel          the earlier one-tile-per-CTA kernel, kept as the ablation baseline
//                                 (T2H_LINEAR_NONPERSISTENT=1)
//   wgrad_x3_kernel, wgrad_reduce_kernel   weight / bias gradient (MN-major operands, row split + fixed-order sum)
//   absmax_kernel, add_absmax_kernel, split_*_kernel, colsum kernels   operand maxima, weight splits, bias gradient
// The stages of every GEMM are connected by mbarrier pipelines (TMA -> split -> MMA -> slot free; MMA -> epilogue).
#include "t2h_common.cuh"
#include "t2h_tc.cuh"
#include <cstdlib>

namespace t2h {
namespace gemm {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // fp32 elements = one 128-byte swizzle row
constexpr int UMMA_K = 8;     // tf32
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
constexpr int kThreads = 192;
// every kernel keeps its mbarriers in a 256-byte block (at most 26 of 32 words used); the word that receives the
// tensor-memory base address from tcgen05.alloc sits at its end, 16-byte aligned and away from the barriers that
// thread 0 initialises at the same time (keeps compute-sanitizer's racecheck quiet about that pair of accesses)
constexpr uint32_t kTmemSlotOffset = 240;

struct LinearArgs {
  int64_t rows;
  int n_out;
  int k_chunks;    // total 32-wide K chunks (source 1 then source 2)
  int k1_chunks;   // chunks taken from the first x source
  int relu_in;
  const float* bias;       // [n_out] or nullptr
  const float* mask;       // [rows, ld_mask] or nullptr: result zeroed where mask <= 0
  int64_t ld_mask;
  const float* residual;   // [rows, ld_res] or nullptr (may alias out)
  int64_t ld_res;
  float* out;              // [rows, ld_out]
  int64_t ld_out;
  int aux_kind;            // which of mask (1) / residual (2) is staged through shared memory by TMA (0: none)
  // 3x3 convolution as implicit GEMM over channels-last planes (B, H, W, C): a 128-row tile is a
  // CONV_TH x CONV_TW pixel patch, K chunk kc = (tap, 32-channel slice), loaded by 4-D TMA from the
  // tap-shifted coordinates (zero fill outside the plane = padding 1)
  int conv;                // 0: plain rows, 1: 3x3 convolution
  int tiles_x;             // patches per plane row  (W / CONV_TW)
  int tiles_per_img;       // patches per plane      ((H / CONV_TH) * tiles_x)
  int cin_chunks;          // Cin / 32
  // fp16x3 mode: device words holding the bit pattern of max|x| (pre-ReLU) and max|w| (t2h_absmax); both
  // operands are scaled by a power of two that brings the maximum just below 2^15 before the fp16 split
  const uint32_t* x_absmax;
  const uint32_t* w_absmax;
  uint32_t* out_absmax;    // nullable: max |out| is merged into this word (operand scale of the next fp16 GEMM)
};
constexpr int CONV_TW = 16, CONV_TH = 8;  // 16 x 8 pixels = 128 GEMM rows

// A_TMEM: the split x operand is written to tensor memory (tcgen05.st) and consumed from there, so the
// three MMAs of a k-step only read the WEIGHT tiles from shared memory.  In SS mode the 128x256 tile is
// shared-memory-bandwidth bound (each MMA re-reads 4 KB of A and 8 KB of B per 134 cycles).
// STAGES_: pipeline depth.  Narrow layers have only 1-2 K chunks; with a single stage a CTA needs 40-48 KB
// of shared memory and 4-5 CTAs share an SM, which hides the serial load -> split -> MMA -> store chain
// of one CTA behind the others (these layers are HBM-bound).
template <int BLOCK_N, bool A_TMEM, int STAGES_>
struct Smem {
  static constexpr int W_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = (A_TMEM ? 1 : 2) * A_BYTES + 2 * W_BYTES;
  static constexpr int STAGES = STAGES_;  // A_TMEM/128 with 2 stages: 96 KB + 256 TMEM cols -> 2 CTAs per SM
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
  // TMEM columns: accumulator + (hi, lo) x 32 columns per stage when A lives in TMEM
  static constexpr int TMEM_USED = (BLOCK_N < 32 ? 32 : BLOCK_N) + (A_TMEM ? STAGES * 64 : 0);
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : (TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512)));
  static constexpr int W_OFF = (A_TMEM ? 1 : 2) * A_BYTES;
};

template <int BLOCK_N, bool A_TMEM, int STAGES_>
__global__ void __launch_bounds__(kThreads)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                     const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo,
                     const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_aux,
                     const LinearArgs p) {
  using S = Smem<BLOCK_N, A_TMEM, STAGES_>;
  constexpr int STAGES = S::STAGES;
  constexpr int W_OFF = S::W_OFF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * S::STAGE_BYTES;
  auto full_tma = [&](int s) { return bars + 8u * s; };
  auto full_ab = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  const uint32_t tmem_full = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bars + kTmemSlotOffset;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * S::STAGE_BYTES + kTmemSlotOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1-D grid, N-tile fastest: the CTAs that share an x tile are co-scheduled, so the tile is fetched
  // from DRAM once and re-read from L2 (an M-fastest raster re-read x from DRAM once per N-tile)
  const int n_tiles = (p.n_out + BLOCK_N - 1) / BLOCK_N;
  const int m_idx = (int)(blockIdx.x / n_tiles);
  const int m0 = m_idx * BLOCK_M;
  const int n0 = (int)(blockIdx.x % n_tiles) * BLOCK_N;
  int img = 0, px0 = 0, py0 = 0;  // conv mode: image and top-left pixel of this patch
  if (p.conv) {
    img = m_idx / p.tiles_per_img;
    const int rem = m_idx - img * p.tiles_per_img;
    py0 = (rem / p.tiles_x) * CONV_TH;
    px0 = (rem % p.tiles_x) * CONV_TW;
  }
  constexpr uint32_t A_TMEM_COL0 = (BLOCK_N < 32 ? 32 : BLOCK_N);  // A stages follow the accumulator columns

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_tma(s), 1);
      mbar_init(full_ab(s), 128);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(bars + 8u * (3 * STAGES + 2), 1);  // aux (mask / residual) tile landed
    fence_barrier_init();
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_whi);
    tma_prefetch_desc(&tm_wlo);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        const uint32_t stage = base + s * S::STAGE_BYTES;
        mbar_arrive_expect_tx(full_tma(s), A_BYTES + 2 * S::W_BYTES);
        if (p.conv) {
          const int tap = kc / p.cin_chunks, cc = kc - tap * p.cin_chunks;
          tma_load_4d(stage, &tm_x1, full_tma(s), cc * BLOCK_K, px0 + tap % 3 - 1, py0 + tap / 3 - 1, img);
        } else if (kc < p.k1_chunks) tma_load_2d(stage, &tm_x1, full_tma(s), kc * BLOCK_K, m0);
        else                  tma_load_2d(stage, &tm_x2, full_tma(s), (kc - p.k1_chunks) * BLOCK_K, m0);
        tma_load_2d(stage + W_OFF, &tm_whi, full_tma(s), kc * BLOCK_K, n0);
        tma_load_2d(stage + W_OFF + S::W_BYTES, &tm_wlo, full_tma(s), kc * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BLOCK_M, BLOCK_N, 0, 0);
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(full_tma(s), ph);
        mbar_wait(full_ab(s), ph);
        tc_fence_after();
        const uint32_t stage = base + s * S::STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint32_t koff = k * UMMA_K * 4;  // 32 bytes along the swizzled 128-byte row
          const uint64_t b_hi = make_smem_desc(stage + W_OFF + koff, 16, 1024);
          const uint64_t b_lo = make_smem_desc(stage + W_OFF + S::W_BYTES + koff, 16, 1024);
          if (A_TMEM) {
            const uint32_t a_hi = tmem_d + A_TMEM_COL0 + s * 64 + k * UMMA_K;
            const uint32_t a_lo = a_hi + 32;
            mma_tf32_ts(tmem_d, a_lo, b_hi, idesc, (kc | k) != 0);
            mma_tf32_ts(tmem_d, a_hi, b_lo, idesc, 1);
            mma_tf32_ts(tmem_d, a_hi, b_hi, idesc, 1);
          } else {
            const uint64_t a_hi = make_smem_desc(stage + koff, 16, 1024);
            const uint64_t a_lo = make_smem_desc(stage + A_BYTES + koff, 16, 1024);
            mma_tf32(tmem_d, a_lo, b_hi, idesc, (kc | k) != 0);
            mma_tf32(tmem_d, a_hi, b_lo, idesc, 1);
            mma_tf32(tmem_d, a_hi, b_hi, idesc, 1);
          }
        }
        mma_commit(empty(s));  // slot reusable once these MMAs have read their operands
      }
      mma_commit(tmem_full);
    }
  } else {
    // ---- operand transform: thread t owns row t of the 128 x 32 chunk ------------------------
    // (a warp may only touch TMEM lanes [32*(warp%4), +32), so rows follow the same quarters)
    const int t = (warp & 3) * 32 + lane;
    for (int kc = 0; kc < p.k_chunks; ++kc) {
      const int s = kc % STAGES;
      const uint32_t ph = (kc / STAGES) & 1;
      mbar_wait(full_tma(s), ph);
      float* hi_row = reinterpret_cast<float*>(base_ptr + s * S::STAGE_BYTES + t * 128);
      if (A_TMEM) {
        // TMA placed logical 16-byte chunk j of row t at chunk j ^ (t & 7) (SWIZZLE_128B)
        float hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = *reinterpret_cast<const float4*>(hi_row + ((j ^ (t & 7)) * 4));
          if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          split_tf32(v.x, hi[4 * j], lo[4 * j]); split_tf32(v.y, hi[4 * j + 1], lo[4 * j + 1]);
          split_tf32(v.z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(v.w, hi[4 * j + 3], lo[4 * j + 3]);
        }
        const uint32_t a_dst = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + A_TMEM_COL0 + s * 64;
        tmem_st_32x32(a_dst, hi);
        tmem_st_32x32(a_dst + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(full_ab(s));
        continue;
      }
      float* lo_row = reinterpret_cast<float*>(base_ptr + s * S::STAGE_BYTES + A_BYTES + t * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = ((j + t) & 7) * 4;  // rotate chunks across threads: conflict-free 16-byte accesses
        float4 v = *reinterpret_cast<float4*>(hi_row + c);
        if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi_row + c) = h;
        *reinterpret_cast<float4*>(lo_row + c) = l;
      }
      fence_proxy_async_smem();
      mbar_arrive(full_ab(s));
    }
    // ---- epilogue ------------------------------------------------------------------------------
    // TMEM -> registers -> bias / ReLU-mask / residual -> shared memory (128-byte swizzled rows) -> one
    // TMA store per 32-column block: global writes (and the mask / residual reads, which arrive by
    // TMA into the same staging rows) are full coalesced lines instead of 16-byte pieces per thread.
    // The pipeline stages are free by now (all TMA loads consumed, all MMAs complete) and are reused.
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int64_t row = (int64_t)m0 + t;
    const bool row_ok = row < p.rows;
    const uint32_t aux_bar = bars + 8u * (3 * STAGES + 2);
    const int n_blocks = min(BLOCK_N / 32, (p.n_out - n0 + 31) / 32);
    static_assert((BLOCK_N / 32) * A_BYTES <= STAGES * S::STAGE_BYTES, "staging does not fit");
    if (p.aux_kind) {
      if (t == 0) {
        mbar_arrive_expect_tx(aux_bar, (uint32_t)n_blocks * A_BYTES);
        const CUtensorMap* aux = &tm_aux;
        for (int cb = 0; cb < n_blocks; ++cb) {
          if (p.conv) tma_load_4d(base + cb * A_BYTES, aux, aux_bar, n0 + cb * 32, px0, py0, img);
          else tma_load_2d(base + cb * A_BYTES, aux, aux_bar, n0 + cb * 32, m0);
        }
      }
      mbar_wait(aux_bar, 0);
    }
#pragma unroll 1
    for (int cb = 0; cb < n_blocks; ++cb) {
      float v[32];
      __syncwarp();
      tmem_ld_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), v);
      float* srow = reinterpret_cast<float*>(base_ptr + cb * A_BYTES + t * 128);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int n = n0 + cb * 32 + g * 4;
        float4* slot = reinterpret_cast<float4*>(srow + ((g ^ (t & 7)) * 4));  // SWIZZLE_128B placement
        float4 o = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        if (n < p.n_out) {
          if (p.bias) {
            const float4 b = ld4(p.bias + n);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (p.mask) {
            float4 m;
            if (p.aux_kind == 1) m = *slot;
            else m = row_ok ? ld4(p.mask + row * p.ld_mask + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f; o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
          }
          if (p.residual) {
            float4 r;
            if (p.aux_kind == 2) r = *slot;
            else r = row_ok ? *reinterpret_cast<const float4*>(p.residual + row * p.ld_res + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
          }
        }
        *slot = o;
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
    if (t == 0) {
      for (int cb = 0; cb < n_blocks; ++cb) {
        if (p.conv) tma_store_4d(&tm_out, base + cb * A_BYTES, n0 + cb * 32, px0, py0, img);
        else tma_store_2d(&tm_out, base + cb * A_BYTES, n0 + cb * 32, m0);
      }
      tma_store_commit_and_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, S::TMEM_COLS);
}


// ---- persistent variant (wide layers) ----------------------------------------------------------------
// One CTA per SM walks the tile list (N-tile fastest, stride gridDim.x).  Roles: warp 0 TMA producer,
// warp 1 MMA issuer, warps 2-5 operand transform (split x -> TMEM), warps 6-13 epilogue.  The fp32
// accumulator is DOUBLE-BUFFERED in tensor memory (2 x 128 columns) and the stage ring runs across tile
// boundaries, so tile i+1 is loaded, split and multiplied while the epilogue warps drain, post-process and
// TMA-store tile i: the ~6 us of per-tile prologue/epilogue that the one-tile-per-CTA kernel exposes
// (TMEM allocation, barrier setup, first-load latency, store drain) is paid once per SM or hidden.
// TMEM: 2 x 128 (accumulators) + 4 x 64 (split A stages) = all 512 columns.
constexpr int P_THREADS = 448;  // producer, MMA, 4 transform warps, 2 x 4 epilogue warps (144 registers per thread)
// F16 = false (3xTF32): a stage is one 32-wide K chunk (x fp32 16 KB | w hi 16 KB | w lo 16 KB), 4 stages.
// F16 = true (3xFP16): a stage is 64 K elements (two fp32 x boxes of 16 KB | w hi fp16 16 KB | w lo fp16 16 KB),
//   3 stages.  The split x operand is packed two fp16 per 32-bit TMEM column, so a stage again occupies
//   64 columns (32 hi + 32 lo).
// BLOCK_N = 128 / 64 / 32 output columns per tile; narrower tiles have smaller weight stages and run a deeper ring
// (the narrow layers are HBM-bound: bytes in flight per SM matter, not MMA issue).
// WSTEPS > 0 (fp16, 128-wide tiles, K <= 64 * WSTEPS): WEIGHT-RESIDENT variant.  With a grid that is a multiple
// of the number of N-tiles the raster `tile += gridDim.x` keeps every CTA on ONE N-tile, so its weight tile
// (K x 128 fp16 hi + lo, <= 128 KB) is loaded once and stays in shared memory; the stage ring then carries the
// x chunks only, which halves the L2 -> SM tile traffic that bounds these layers.
// PAIR (fp16, 128-wide tiles): two CTAs of a cluster (the two SMs of a TPC) compute a 256 x 128 tile with
// tcgen05.mma.cta_group::2 -- each CTA transforms its own 128 rows of x into its own tensor memory, stages only
// HALF of the weight tile (64 of the 128 output columns; the MMA reads both halves) and drains its own
// accumulator.  48 KB instead of 64 KB per step and SM, so four stages fit: fewer bytes per MMA cycle and a
// deeper load pipeline, which is what bounds the single-CTA kernel.
// DS (CTA pairs): TWO epilogue staging slots per group and three load stages.  With one slot a group's blocks run
// strictly one after the other -- wait until the previous store has been read out of the slot, load the mask /
// residual block into it (a full memory latency), post-process, store -- which bounds the layers with K <= 512
// (a tile's MMAs last 2-4 stages).  With two slots both mask / residual loads of a tile are issued before the
// accumulator is waited for, and a store drains while the next block is processed.
template <bool F16, int BLOCK_N, int WSTEPS = 0, bool PAIR = false, bool DS = false>
struct PSmem {
  static constexpr bool WRES = WSTEPS > 0;
  // WIDE (CTA pairs, 256-wide tiles): a 256 x 256 tile per pair.  Both accumulators (2 x 256 columns) fill the
  // tensor memory, so the split x operand stays in SHARED memory: the transform warps overwrite each 128-byte fp32
  // row of an x box (32 k) with [hi: 32 fp16 | lo: 32 fp16] in place -- still a K-major SWIZZLE_128B tile, whose
  // k-steps 0,1 are the hi and 2,3 the lo halves -- and the MMAs read A through shared-memory descriptors.  One
  // conversion and one L2 -> SM transfer of x now serves 256 output columns instead of 128.
  static constexpr bool WIDE = PAIR && BLOCK_N == 256;
  static constexpr int STAGES = WIDE ? (DS ? 2 : 3) : PAIR ? (DS ? 3 : 4) : (WRES ? (WSTEPS <= 2 ? 4 : 2) : (F16 ? (BLOCK_N == 128 ? 3 : 4) : (BLOCK_N == 128 ? 4 : 6)));
  static constexpr int X_BYTES = (F16 ? 2 : 1) * A_BYTES;
  static constexpr int W_BYTES = (F16 ? BLOCK_N * 64 * 2 : BLOCK_N * BLOCK_K * 4) / (PAIR ? 2 : 1);
  static constexpr int STAGE_BYTES = WRES ? X_BYTES : X_BYTES + 2 * W_BYTES;   // x raw (| w hi | w lo)
  static constexpr int WREGION_BYTES = WSTEPS * 2 * W_BYTES;       // resident weight: per k-step (w hi | w lo)
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES + WREGION_BYTES;
  static constexpr int STAGING_BYTES = (DS ? 4 : 2) * A_BYTES;     // epilogue staging: 32-column blocks, one (DS: two) per group
  static constexpr int TOTAL = STAGING_OFF + STAGING_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int A_COL0 = 2 * BLOCK_N;
  static_assert(!WRES || (F16 && BLOCK_N == 128), "weight-resident: fp16, 128-wide tiles");
  static_assert(!PAIR || (F16 && (BLOCK_N == 128 || BLOCK_N == 256) && !WRES), "CTA pairs: fp16, 128- or 256-wide tiles");
  static_assert(!DS || PAIR, "double staging slots: CTA pairs");
  static_assert(WIDE ? A_COL0 == 512 : A_COL0 + STAGES * 64 <= 512, "tensor memory: two accumulators (+ the split x stages)");
  static_assert(TOTAL <= 227 * 1024, "shared memory");
};

// power-of-two operand scale from the bit pattern of max|v|: 2^e with e = 14 - floor(log2 max), so that the
// scaled maximum lies in [2^14, 2^15) (fp16 overflows at 65504); |e| <= 63 (maxima between 2^-49 and 2^77)
__device__ __forceinline__ int f16_scale_exp(uint32_t absmax_bits) {
  const int biased = (int)((absmax_bits >> 23) & 0xFF);
  if (biased == 0) return 0;  // zero / denormal maximum: nothing to scale
  int e = 14 - (biased - 127);
  return e > 63 ? 63 : (e < -63 ? -63 : e);  // two such scales multiply to a representable float
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }

template <bool F16, int BLOCK_N, int WSTEPS, bool PAIR, bool DS>
__global__ void __launch_bounds__(P_THREADS, 1)
linear_x3_persistent_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
                                const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo,
                                const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_aux,
                                const LinearArgs p, const int64_t n_tiles_total) {
  using S = PSmem<F16, BLOCK_N, WSTEPS, PAIR, DS>;
  // PAIR: rank of this CTA in its pair; the pair (not the CTA) walks the list of 256-row tiles
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int64_t vblock = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t vgrid = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  constexpr int STAGES = S::STAGES;
  constexpr bool WRES = S::WRES;
  // pipeline steps per tile: 32-wide chunks (tf32) or pairs of them (fp16)
  const int n_steps = F16 ? (p.k_chunks + 1) / 2 : p.k_chunks;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t wregion = base + STAGES * S::STAGE_BYTES;  // (weight-resident variant)
  const uint32_t staging = base + S::STAGING_OFF;
  uint8_t* staging_ptr = base_ptr + S::STAGING_OFF;
  const uint32_t bars = staging + S::STAGING_BYTES;
  auto full_tma = [&](int s) { return bars + 8u * s; };
  auto full_ab = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto acc_full = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto acc_empty = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  auto aux_bar = [&](int g) { return bars + 8u * (3 * STAGES + 4 + g); };
  const uint32_t w_full = bars + 8u * (3 * STAGES + 7);
  auto w_pair = [&](int s) { return bars + 8u * (3 * STAGES + 8 + s); };  // PAIR: both weight halves landed (leader's)
  auto aux_bar2 = [&](int g, int j) { return bars + 8u * (4 * STAGES + 8 + 2 * g + j); };  // DS: per group and slot
  const uint32_t tmem_slot = bars + kTmemSlotOffset;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(staging_ptr + S::STAGING_BYTES + kTmemSlotOffset);
  // epilogue threads that hand an accumulator back per tile: both groups, or (32-wide tiles) the one that owns the tile
  constexpr uint32_t EPI_ARRIVALS = BLOCK_N == 32 ? 4 : 8;  // one arrival per epilogue warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.n_out + BLOCK_N - 1) / BLOCK_N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_tma(s), 1);
      mbar_init(full_ab(s), PAIR ? 8 : 4);      // one arrival per transform warp (PAIR: of both CTAs, on the leader's)
      mbar_init(empty(s), 1);
      if (PAIR) mbar_init(w_pair(s), 2);        // the two producers (the leader's arrival carries the byte count)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), PAIR ? 2 * EPI_ARRIVALS : EPI_ARRIVALS);  // PAIR: both CTAs' epilogues, on the leader's
      mbar_init(aux_bar(a), 1);
      if (DS) { mbar_init(aux_bar2(a, 0), 1); mbar_init(aux_bar2(a, 1), 1); }
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_whi);
    tma_prefetch_desc(&tm_wlo);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, S::TMEM_COLS);
    else tmem_alloc(tmem_slot, S::TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // tile -> coordinates
  auto decode = [&](int64_t tile, int& m0, int& n0, int& img, int& px0, int& py0) {
    const int m_idx = PAIR ? 2 * (int)(tile / n_tiles) + (int)rank : (int)(tile / n_tiles);
    m0 = m_idx * BLOCK_M;
    n0 = (int)(tile % n_tiles) * BLOCK_N;
    img = px0 = py0 = 0;
    if (p.conv) {
      img = m_idx / p.tiles_per_img;
      const int rem = m_idx - img * p.tiles_per_img;
      py0 = (rem / p.tiles_x) * CONV_TH;
      px0 = (rem % p.tiles_x) * CONV_TW;
    }
  };

  if (warp == 0) {
    {  // TMA producer: the warp runs the loop, one elected lane issues (see tc::elect_one)
      if (WRES && (int64_t)blockIdx.x < n_tiles_total) {
        // the CTA's one weight tile: every k-step's (hi | lo) pair, once
        if (elect_one()) {
          const int n0w = (int)(blockIdx.x % n_tiles) * BLOCK_N;
          mbar_arrive_expect_tx(w_full, (uint32_t)n_steps * 2 * S::W_BYTES);
          for (int kc = 0; kc < n_steps; ++kc) {
            tma_load_2d(wregion + kc * 2 * S::W_BYTES, &tm_whi, w_full, kc * 64, n0w);
            tma_load_2d(wregion + kc * 2 * S::W_BYTES + S::W_BYTES, &tm_wlo, w_full, kc * 64, n0w);
          }
        }
        __syncwarp();
      }
      int64_t it = 0;  // running chunk counter across tiles
      for (int64_t tile = vblock; tile < n_tiles_total; tile += vgrid) {
        int m0, n0, img, px0, py0;
        decode(tile, m0, n0, img, px0, py0);
        for (int kc = 0; kc < n_steps; ++kc, ++it) {
          const int s = (int)(it % STAGES);
          const uint32_t ph = (uint32_t)((it / STAGES) & 1);
          mbar_wait(empty(s), ph ^ 1);
          if (elect_one()) {
          const uint32_t stage = base + s * S::STAGE_BYTES;
          mbar_arrive_expect_tx(full_tma(s), S::X_BYTES + ((WRES || PAIR) ? 0 : 2 * S::W_BYTES));
          if (PAIR) {
            // this CTA's 64 weight rows (hi | lo) of the step; the bytes of both CTAs are counted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(w_pair(s), 4 * S::W_BYTES);
            else mbar_arrive_remote(w_pair(s), 0);
            tma_load_2d_pair(stage + S::X_BYTES, &tm_whi, w_pair(s), kc * 64, n0 + (int)rank * (BLOCK_N / 2));
            tma_load_2d_pair(stage + S::X_BYTES + S::W_BYTES, &tm_wlo, w_pair(s), kc * 64, n0 + (int)rank * (BLOCK_N / 2));
          }
          auto load_x = [&](uint32_t dst, int c) {  // 32-wide chunk c of the (concatenated / im2col) K axis
            if (p.conv) {
              const int tap = c / p.cin_chunks, cc = c - tap * p.cin_chunks;
              // chunks past the end (odd chunk count in fp16 mode) read outside the channel axis: zero fill
              tma_load_4d(dst, &tm_x1, full_tma(s), cc * BLOCK_K + (tap >= 9 ? (1 << 20) : 0), px0 + tap % 3 - 1, py0 + tap / 3 - 1, img);
            } else if (c < p.k1_chunks) tma_load_2d(dst, &tm_x1, full_tma(s), c * BLOCK_K, m0);
            else                        tma_load_2d(dst, &tm_x2, full_tma(s), (c - p.k1_chunks) * BLOCK_K + (c >= p.k_chunks ? (1 << 20) : 0), m0);
          };
          if (F16) {
            load_x(stage, 2 * kc);
            load_x(stage + A_BYTES, 2 * kc + 1);
            if (!WRES && !PAIR) {
              tma_load_2d(stage + S::X_BYTES, &tm_whi, full_tma(s), kc * 64, n0);
              tma_load_2d(stage + S::X_BYTES + S::W_BYTES, &tm_wlo, full_tma(s), kc * 64, n0);
            }
          } else {
            load_x(stage, kc);
            tma_load_2d(stage + S::X_BYTES, &tm_whi, full_tma(s), kc * BLOCK_K, n0);
            tma_load_2d(stage + S::X_BYTES + S::W_BYTES, &tm_wlo, full_tma(s), kc * BLOCK_K, n0);
          }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop (waits included); one elected lane issues the MMAs and their commits
    // (PAIR: only the leader CTA issues; its MMAs drive the tensor cores of both SMs)
    constexpr uint32_t idesc = F16 ? make_idesc_f16(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N, 0, 0) : make_idesc_tf32(BLOCK_M, BLOCK_N, 0, 0);
    if (!PAIR || rank == 0) {
    const uint32_t a_base = tmem_base + S::A_COL0;
    if (WRES && (int64_t)blockIdx.x < n_tiles_total) mbar_wait(w_full, 0);
    int64_t it = 0, local = 0;
    for (int64_t tile = vblock; tile < n_tiles_total; tile += vgrid, ++local) {
      const int acc = (int)(local & 1);
      const uint32_t acc_ph = (uint32_t)((local >> 1) & 1);
      if (PAIR) mbar_wait_cluster(acc_empty(acc), acc_ph ^ 1);
      else mbar_wait(acc_empty(acc), acc_ph ^ 1);  // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      for (int kc = 0; kc < n_steps; ++kc, ++it) {
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
        if (PAIR) {
          mbar_wait_cluster(w_pair(s), ph);
          mbar_wait_cluster(full_ab(s), ph);
        } else {
          mbar_wait(full_tma(s), ph);
          mbar_wait(full_ab(s), ph);
        }
        tc_fence_after();
        if (elect_one()) {
          // every k-step advances 32 bytes along the swizzled 128-byte weight row (8 tf32 or 16 fp16) and
          // 8 TMEM columns of the split x operand; small terms first
          const uint32_t w_tile = WRES ? wregion + kc * 2 * S::W_BYTES : base + s * S::STAGE_BYTES + S::X_BYTES;
          const uint64_t b_hi0 = make_smem_desc(w_tile, 16, 1024);
          const uint64_t b_lo0 = make_smem_desc(w_tile + S::W_BYTES, 16, 1024);
          const uint32_t a_hi0 = a_base + s * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b_hi = b_hi0 + (uint64_t)(k * 2), b_lo = b_lo0 + (uint64_t)(k * 2);  // +32 bytes (>> 4)
            const uint32_t a_hi = a_hi0 + k * 8, a_lo = a_hi + 32;
            if constexpr (S::WIDE) {
              // A from shared memory: box k / 2 of the stage, (hi | lo) halves of its 128-byte rows, 32 bytes per k-step
              const uint64_t sa_hi = make_smem_desc(base + s * S::STAGE_BYTES + (k >> 1) * A_BYTES + (k & 1) * 32, 16, 1024);
              const uint64_t sa_lo = sa_hi + 4;  // + 64 bytes (>> 4)
              mma_f16_ss_pair(tmem_d, sa_lo, b_hi, idesc, (kc | k) != 0);
              mma_f16_ss_pair(tmem_d, sa_hi, b_lo, idesc, 1);
              mma_f16_ss_pair(tmem_d, sa_hi, b_hi, idesc, 1);
            } else if (PAIR) {
              mma_f16_ts_pair(tmem_d, a_lo, b_hi, idesc, (kc | k) != 0);
              mma_f16_ts_pair(tmem_d, a_hi, b_lo, idesc, 1);
              mma_f16_ts_pair(tmem_d, a_hi, b_hi, idesc, 1);
            } else if (F16) {
              mma_f16_ts(tmem_d, a_lo, b_hi, idesc, (kc | k) != 0);
              mma_f16_ts(tmem_d, a_hi, b_lo, idesc, 1);
              mma_f16_ts(tmem_d, a_hi, b_hi, idesc, 1);
            } else {
              mma_tf32_ts(tmem_d, a_lo, b_hi, idesc, (kc | k) != 0);
              mma_tf32_ts(tmem_d, a_hi, b_lo, idesc, 1);
              mma_tf32_ts(tmem_d, a_hi, b_hi, idesc, 1);
            }
          }
          if (PAIR) mma_commit_pair(empty(s));
          else mma_commit(empty(s));
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (PAIR) mma_commit_pair(acc_full(acc));
        else mma_commit(acc_full(acc));
      }
      __syncwarp();
    }
    }
  } else if (warp < 6) {
    // ---- operand transform: row t of every x chunk -> (hi | lo) in TMEM ---------------------------------
    const int t = (warp & 3) * 32 + lane;
    const float sx = F16 ? pow2f(f16_scale_exp(*p.x_absmax)) : 1.f;
    int64_t it = 0;
    for (int64_t tile = vblock; tile < n_tiles_total; tile += vgrid) {
      for (int kc = 0; kc < n_steps; ++kc, ++it) {
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
        mbar_wait(full_tma(s), ph);
        const uint32_t a_dst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + S::A_COL0 + s * 64;
        if constexpr (S::WIDE) {
#pragma unroll
          for (int box = 0; box < 2; ++box) {
            uint8_t* row = base_ptr + s * S::STAGE_BYTES + box * A_BYTES + t * 128;
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] = *reinterpret_cast<const float4*>(row + ((j ^ (t & 7)) << 4));
              if (p.relu_in) { v[j].x = fmaxf(v[j].x, 0.f); v[j].y = fmaxf(v[j].y, 0.f); v[j].z = fmaxf(v[j].z, 0.f); v[j].w = fmaxf(v[j].w, 0.f); }
              v[j].x *= sx; v[j].y *= sx; v[j].z *= sx; v[j].w *= sx;
            }
            uint4 h[4], l[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              split_f16x2(v[2 * c].x, v[2 * c].y, h[c].x, l[c].x);
              split_f16x2(v[2 * c].z, v[2 * c].w, h[c].y, l[c].y);
              split_f16x2(v[2 * c + 1].x, v[2 * c + 1].y, h[c].z, l[c].z);
              split_f16x2(v[2 * c + 1].z, v[2 * c + 1].w, h[c].w, l[c].w);
            }
            // the whole fp32 row is in registers: overwrite it with 16-byte chunks 0-3 = hi, 4-7 = lo (same swizzle)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              *reinterpret_cast<uint4*>(row + ((c ^ (t & 7)) << 4)) = h[c];
              *reinterpret_cast<uint4*>(row + (((4 + c) ^ (t & 7)) << 4)) = l[c];
            }
          }
          fence_proxy_async_smem();
        } else if (F16) {
#pragma unroll
          for (int box = 0; box < 2; ++box) {
            const float* x_row = reinterpret_cast<const float*>(base_ptr + s * S::STAGE_BYTES + box * A_BYTES + t * 128);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 v = *reinterpret_cast<const float4*>(x_row + ((j ^ (t & 7)) * 4));
              if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              v.x *= sx; v.y *= sx; v.z *= sx; v.w *= sx;
              split_f16x2(v.x, v.y, hi[2 * j], lo[2 * j]);
              split_f16x2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
            }
            tmem_st_32x16(a_dst + box * 16, hi);
            tmem_st_32x16(a_dst + 32 + box * 16, lo);
          }
        } else {
          const float* x_row = reinterpret_cast<const float*>(base_ptr + s * S::STAGE_BYTES + t * 128);
          float hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = *reinterpret_cast<const float4*>(x_row + ((j ^ (t & 7)) * 4));
            if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            split_tf32(v.x, hi[4 * j], lo[4 * j]); split_tf32(v.y, hi[4 * j + 1], lo[4 * j + 1]);
            split_tf32(v.z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(v.w, hi[4 * j + 3], lo[4 * j + 3]);
          }
          tmem_st_32x32(a_dst, hi);
          tmem_st_32x32(a_dst + 32, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // one (possibly remote) arrival per warp
          if (PAIR && rank != 0) mbar_arrive_remote(full_ab(s), 0);
          else mbar_arrive(full_ab(s));
        }
      }
    }
  } else {
    // ---- epilogue warps 6-13: drain accumulator `acc`, release it, post-process, TMA store --------------
    // Two independent groups of four warps (one warp per TMEM lane quarter and SM sub-partition each), each
    // with its own 16 KB staging slot, named barrier, TMA-store leader and aux barrier.  Group eh handles the
    // 32-column blocks eh, eh + 2 of a tile; with 32-wide tiles the groups take alternate tiles.  A group only
    // waits for ITS previous store to have been read out of shared memory right before it refills the slot.
    const int quarter = warp & 3;  // warps 6..13 -> TMEM lane quarters 2,3,0,1,2,3,0,1
    const int eh = (warp - 6) >> 2;
    const int t = quarter * 32 + lane;
    const bool leader = ((warp - 6) & 3) == 0 && lane == 0;
    const uint32_t group_staging = staging + eh * (DS ? 2 : 1) * A_BYTES;
    float* group_srow = reinterpret_cast<float*>(staging_ptr + eh * (DS ? 2 : 1) * A_BYTES + t * 128);
    // fp16 mode: undo the two power-of-two operand scales (|exponent| <= 63 each, so the product is a float)
    const float inv = F16 ? pow2f(-(f16_scale_exp(*p.x_absmax) + f16_scale_exp(*p.w_absmax))) : 1.f;
    const int epi_mode = (p.mask && p.residual) || (p.aux_kind == 0 && (p.mask || p.residual)) ? 3 : p.aux_kind;
    uint32_t aux_phase = 0;
    bool prev_even = false;  // DS: the group's latest store came from slot 1 (an even number of blocks in its last tile)
    float out_max = 0.f;
    int64_t local = 0;
    for (int64_t tile = vblock; tile < n_tiles_total; tile += vgrid, ++local) {
      if (BLOCK_N == 32 && (int)(local & 1) != eh) continue;
      int m0, n0, img, px0, py0;
      decode(tile, m0, n0, img, px0, py0);
      const int acc = (int)(local & 1);
      const uint32_t acc_ph = (uint32_t)((local >> 1) & 1);
      const int n_blocks = min(BLOCK_N / 32, (p.n_out - n0 + 31) / 32);
      const int64_t row = (int64_t)m0 + t;
      const bool row_ok = row < p.rows;
      const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
      bool have_acc = false;
      // DS: the group's blocks of this tile (2 of a 128-wide, 4 of a 256-wide tile) alternate between its two slots
      const int nb_grp = DS && n_blocks > eh ? (n_blocks - eh + 1) >> 1 : 0;
      auto load_aux = [&](int b) {  // (leader) mask / residual block b of the group -> slot b & 1
        const uint32_t dst = group_staging + (b & 1) * A_BYTES, bar = aux_bar2(eh, b & 1);
        mbar_arrive_expect_tx(bar, A_BYTES);
        if (p.conv) tma_load_4d(dst, &tm_aux, bar, n0 + (eh + 2 * b) * 32, px0, py0, img);
        else tma_load_2d(dst, &tm_aux, bar, n0 + (eh + 2 * b) * 32, m0);
      };
      if (DS && p.aux_kind && leader && nb_grp > 0) {
        // the first two mask / residual blocks of the tile go out before the accumulator is waited for; a slot is
        // free once the store that last used it has been read (slot 0: all but the latest store, if that was slot 1's)
        if (prev_even) tma_store_wait_read_but_one(); else tma_store_wait_read();
        load_aux(0);
        if (nb_grp > 1) {
          tma_store_wait_read();
          load_aux(1);
        }
      }
#pragma unroll 1
      for (int cb = (BLOCK_N == 32 ? 0 : eh); cb < n_blocks; cb += 2) {
        const int blk = (cb - eh) >> 1;  // (DS) index of the block within the group's share of the tile
        const int slot = DS ? blk & 1 : 0;
        const uint32_t my_staging = group_staging + slot * A_BYTES;
        float* srow = group_srow + slot * (A_BYTES / 4);
        if (DS && p.aux_kind) {
          mbar_wait(aux_bar2(eh, slot), (aux_phase >> slot) & 1u);
          aux_phase ^= 1u << slot;
        } else {
          // the slot is free once the store that last used it has been read
          if (leader) {
            if (DS && (blk > 0 || prev_even)) tma_store_wait_read_but_one();
            else tma_store_wait_read();
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + eh) : "memory");
          if (p.aux_kind) {
            if (leader) {
              mbar_arrive_expect_tx(aux_bar(eh), A_BYTES);
              if (p.conv) tma_load_4d(my_staging, &tm_aux, aux_bar(eh), n0 + cb * 32, px0, py0, img);
              else tma_load_2d(my_staging, &tm_aux, aux_bar(eh), n0 + cb * 32, m0);
            }
            mbar_wait(aux_bar(eh), aux_phase);
            aux_phase ^= 1;
          }
        }
        if (!have_acc) {
          mbar_wait(acc_full(acc), acc_ph);
          tc_fence_after();
          have_acc = true;
        }
        const int nb = n0 + cb * 32;
        // the bias slice first (one broadcast line per load, in flight while tensor memory is read); columns
        // past n_out are clipped by the TMA store, so their index is only clamped to stay in bounds
        float4 b[8];
        if (p.bias) {
#pragma unroll
          for (int g = 0; g < 8; ++g) b[g] = ld4(p.bias + min(nb + g * 4, p.n_out - 4));
        }
        float v[32];
        tmem_ld_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), v);
        if (cb + 2 >= n_blocks) {
          // the group's last read of this accumulator (tcgen05.ld has been waited for)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_remote(acc_empty(acc), 0);
            else mbar_arrive(acc_empty(acc));
          }
        }
        if (F16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= inv;
        }
        if (p.bias) {
#pragma unroll
          for (int g = 0; g < 8; ++g) { v[4 * g] += b[g].x; v[4 * g + 1] += b[g].y; v[4 * g + 2] += b[g].z; v[4 * g + 3] += b[g].w; }
        }
        if (epi_mode == 1) {         // ReLU mask staged by TMA (input-gradient GEMM)
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 m = *reinterpret_cast<const float4*>(srow + ((g ^ (t & 7)) * 4));
            v[4 * g] = m.x > 0.f ? v[4 * g] : 0.f; v[4 * g + 1] = m.y > 0.f ? v[4 * g + 1] : 0.f;
            v[4 * g + 2] = m.z > 0.f ? v[4 * g + 2] : 0.f; v[4 * g + 3] = m.w > 0.f ? v[4 * g + 3] : 0.f;
          }
        } else if (epi_mode == 2) {  // residual staged by TMA
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 r = *reinterpret_cast<const float4*>(srow + ((g ^ (t & 7)) * 4));
            v[4 * g] += r.x; v[4 * g + 1] += r.y; v[4 * g + 2] += r.z; v[4 * g + 3] += r.w;
          }
        } else if (epi_mode == 3) {  // general case: mask and residual together, read from global memory
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int n = min(nb + g * 4, p.n_out - 4);  // clipped columns: any in-bounds address
            if (p.mask) {
              float4 m;
              if (p.aux_kind == 1) m = *reinterpret_cast<const float4*>(srow + ((g ^ (t & 7)) * 4));
              else m = row_ok ? ld4(p.mask + row * p.ld_mask + n) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * g] = m.x > 0.f ? v[4 * g] : 0.f; v[4 * g + 1] = m.y > 0.f ? v[4 * g + 1] : 0.f;
              v[4 * g + 2] = m.z > 0.f ? v[4 * g + 2] : 0.f; v[4 * g + 3] = m.w > 0.f ? v[4 * g + 3] : 0.f;
            }
            if (p.residual) {
              const float4 r = row_ok ? *reinterpret_cast<const float4*>(p.residual + row * p.ld_res + n) : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * g] += r.x; v[4 * g + 1] += r.y; v[4 * g + 2] += r.z; v[4 * g + 3] += r.w;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(srow + ((g ^ (t & 7)) * 4)) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        if (p.out_absmax && row_ok) {  // rows / columns outside the matrix hold bias-only garbage: skip them
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (nb + g * 4 < p.n_out)
              out_max = fmaxf(fmaxf(out_max, fmaxf(fabsf(v[4 * g]), fabsf(v[4 * g + 1]))), fmaxf(fabsf(v[4 * g + 2]), fabsf(v[4 * g + 3])));
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eh) : "memory");
        if (leader) {
          if (p.conv) tma_store_4d(&tm_out, my_staging, n0 + cb * 32, px0, py0, img);
          else tma_store_2d(&tm_out, my_staging, n0 + cb * 32, m0);
          tma_store_commit();
          if (DS && p.aux_kind && blk + 2 < nb_grp) {  // 256-wide tiles: the slot's next mask / residual block
            tma_store_wait_read();
            load_aux(blk + 2);
          }
        }
      }
      if (DS && nb_grp > 0) prev_even = (nb_grp & 1) == 0;
      if (BLOCK_N != 32 && !have_acc) {
        // a group without a block in this (narrow last) tile still takes part in the accumulator hand-over, in step
        mbar_wait(acc_full(acc), acc_ph);
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_remote(acc_empty(acc), 0);
          else mbar_arrive(acc_empty(acc));
        }
      }
    }
    if (leader) tma_store_wait_all();  // global writes of the last stores complete before the CTA exits
    if (p.out_absmax) {
      uint32_t b = __float_as_uint(out_max);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
      if (lane == 0 && b) atomicMax(p.out_absmax, b);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // nobody's barriers, shared or tensor memory are touched by the peer any more
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, S::TMEM_COLS);
    else tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// ---- weight gradient: dW[n, k] = sum_r g[r, n] * act(x[r, k]),  db[n] = sum_r g[r, n] ---------------
// The reduction runs over the ROWS, so the operands are "MN-major".
//   A = g^T: the row-major tile g[32 rows x 128 n] is TMA-loaded un-swizzled; thread n reads its column
//       (conflict-free), splits it and writes it with tcgen05.st into TMEM lane n -- the transpose is
//       free, the A operand never touches shared memory again, and the same thread accumulates the
//       bias gradient (column sum) on the way.
//   B = x:  row-major tile x[32 rows x BLOCK_N k] TMA-loaded as 32-float boxes with
//       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and consumed with SWIZZLE_128B_BASE32B MN-major descriptors
//       (the only tf32 MN-major layout); hi/lo split element-wise in place.
// Each CTA owns one 128 x BLOCK_N tile of dW for one slice of the rows and writes an fp32 partial;
// partials are summed in a fixed order by wgrad_reduce_kernel (deterministic, no atomics).
constexpr int WG_ROWS = 32;                    // rows per pipeline stage (4 MMA k-steps of 8)
constexpr int WG_GROUP_BYTES = WG_ROWS * 128;  // one 32-float-wide box
constexpr int WG_A_BYTES = WG_ROWS * BLOCK_M * 4;

struct WgradArgs {
  int64_t rows;
  int n_out;       // M extent (columns of g)
  int k_in;        // N extent (columns of x)
  int relu_in;
  int64_t rows_per_split;  // multiple of WG_ROWS
  float* partial;       // [splits, n_out, k_in]
  float* partial_bias;  // [splits, n_out] or nullptr
  // 3x3 convolution: "rows" are pixels taken 32 at a time as 2 x 16 patches of a channels-last plane,
  // the x operand of k-group (tap, channel slice) is the tap-shifted patch (4-D TMA, zero fill = padding)
  int conv;
  int tiles_x;        // W / 16
  int units_per_img;  // (H / 2) * tiles_x
  int cin;
  int a_cols;         // columns of g actually loaded per tile: min(128, n_out rounded up to 32)
  // fp16x3 flavour: device words with the bit patterns of max |g| and max |x| (t2h_absmax)
  const uint32_t* g_absmax;
  const uint32_t* x_absmax;
};

// F16 (BLOCK_N = 128 only): the x tile is TMA-loaded un-swizzled, converted to fp16 hi / lo by the transform
// warps and re-written as two MN-major SWIZZLE_128B operands (64 fp16 = 128 B along k_in per row, 8-row atoms
// 1024 B apart, the two 64-wide groups 4096 B apart); g^T goes to tensor memory as packed fp16 pairs.  A stage
// then needs 2 k-steps of 16 rows (6 MMAs) instead of 4 of 8 (12 MMAs).
template <int BLOCK_N, bool F16 = false>
struct WgSmem {
  static constexpr int B_B = (BLOCK_N / 32) * WG_GROUP_BYTES;    // fp32 x tile (and each of its tf32 splits)
  static constexpr int B16_B = WG_ROWS * BLOCK_N * 2;            // one fp16 operand tile
  static constexpr int STAGE_BYTES = F16 ? WG_A_BYTES + B_B + 2 * B16_B   // g raw | x raw | x hi | x lo (fp16)
                                         : WG_A_BYTES + 2 * B_B;          // g raw | x hi | x lo
  static constexpr int STAGES = BLOCK_N <= 64 ? 3 : 2;           // narrow tiles: 24-32 KB stages, deeper prefetch
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 256 + 1024;
  static constexpr int A_COLS = F16 ? 32 : 64;                   // TMEM columns of (g hi, g lo) per stage
  static constexpr int TMEM_USED = BLOCK_N + STAGES * A_COLS;    // accumulator + split g per stage
  static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512);
};

// producer warp, MMA warp, 4 warps that split g^T into tensor memory + half of the x tile and run the epilogue,
// 4 more warps for the other half of the x tile (the operand conversion is what bounds this kernel)
constexpr int WG_THREADS = 320;
template <int BLOCK_N, bool F16>
__global__ void __launch_bounds__(WG_THREADS)
wgrad_x3_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x, const WgradArgs p) {
  using S = WgSmem<BLOCK_N, F16>;
  constexpr int STAGES = S::STAGES;
  constexpr int B_B = S::B_B;
  static_assert(!F16 || BLOCK_N == 128, "the fp16 flavour is written for 128-wide tiles");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * S::STAGE_BYTES;
  auto full_tma = [&](int s) { return bars + 8u * s; };
  auto full_ab = [&](int s) { return bars + 8u * (STAGES + s); };
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  const uint32_t tmem_full = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bars + kTmemSlotOffset;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * S::STAGE_BYTES + kTmemSlotOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_M;   // rows of dW
  const int k0 = blockIdx.y * BLOCK_N;   // columns of dW
  const int64_t r_begin = (int64_t)blockIdx.z * p.rows_per_split;
  const int64_t r_end = min(r_begin + p.rows_per_split, p.rows);
  const int n_iter = r_end > r_begin ? (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_tma(s), 1);
      mbar_init(full_ab(s), 256);
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    {  // TMA producer: the warp runs the loop, one elected lane issues
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty(s), ph ^ 1);
        if (elect_one()) {
        const uint32_t stage = base + s * S::STAGE_BYTES;
        const int r = (int)(r_begin + (int64_t)it * WG_ROWS);
        mbar_arrive_expect_tx(full_tma(s), (uint32_t)(WG_ROWS * p.a_cols * 4) + B_B);
        if (p.conv) {
          const int u = r >> 5, img = u / p.units_per_img, rem = u - img * p.units_per_img;
          const int y0 = (rem / p.tiles_x) * 2, x0 = (rem % p.tiles_x) * 16;
          tma_load_4d(stage, &tm_g, full_tma(s), n0, x0, y0, img);
#pragma unroll
          for (int gq = 0; gq < BLOCK_N / 32; ++gq) {
            const int k = k0 + gq * 32;
            int tap = k / p.cin, ci0 = k - tap * p.cin;
            if (tap >= 9) { tap = 4; ci0 = p.cin; }  // beyond the 9 taps: out-of-range channel -> zero fill
            tma_load_4d(stage + WG_A_BYTES + gq * WG_GROUP_BYTES, &tm_x, full_tma(s), ci0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
          }
        } else {
          // the split size is a multiple of WG_ROWS, so only the global tail is partial; TMA zero-fills it
          tma_load_2d(stage, &tm_g, full_tma(s), n0, r);
#pragma unroll
          for (int gq = 0; gq < BLOCK_N / 32; ++gq)
            tma_load_2d(stage + WG_A_BYTES + gq * WG_GROUP_BYTES, &tm_x, full_tma(s), k0 + gq * 32, r);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // the whole warp runs the loop; one elected lane issues the MMAs and commits (see tc::elect_one)
    constexpr uint32_t idesc = F16 ? make_idesc_f16(BLOCK_M, BLOCK_N, 0, 1) : make_idesc_tf32(BLOCK_M, BLOCK_N, 0, 1);  // A from TMEM, B MN-major
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(full_tma(s), ph);
      mbar_wait(full_ab(s), ph);
      tc_fence_after();
      if (elect_one()) {
        if (F16) {
          const uint32_t b_hi0 = base + s * S::STAGE_BYTES + WG_A_BYTES + B_B;
          const uint64_t d_hi0 = make_smem_desc(b_hi0, 4096, 1024, kLayoutSW128);
          const uint64_t d_lo0 = make_smem_desc(b_hi0 + S::B16_B, 4096, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < WG_ROWS / 16; ++k) {
            const uint64_t b_hi = d_hi0 + (uint64_t)(k * 128), b_lo = d_lo0 + (uint64_t)(k * 128);  // two 8-row atoms = 2048 bytes
            const uint32_t a_hi = tmem_d + BLOCK_N + s * S::A_COLS + k * 8, a_lo = a_hi + 16;
            mma_f16_ts(tmem_d, a_lo, b_hi, idesc, (it | k) != 0);
            mma_f16_ts(tmem_d, a_hi, b_lo, idesc, 1);
            mma_f16_ts(tmem_d, a_hi, b_hi, idesc, 1);
          }
        } else {
          const uint32_t b_hi0 = base + s * S::STAGE_BYTES + WG_A_BYTES;
          const uint64_t d_hi0 = make_smem_desc(b_hi0, WG_GROUP_BYTES, 512, kLayoutSW128Base32);
          const uint64_t d_lo0 = make_smem_desc(b_hi0 + B_B, WG_GROUP_BYTES, 512, kLayoutSW128Base32);
#pragma unroll
          for (int k = 0; k < WG_ROWS / UMMA_K; ++k) {
            const uint64_t b_hi = d_hi0 + (uint64_t)(k * 64), b_lo = d_lo0 + (uint64_t)(k * 64);  // 8 rows x 128 bytes
            const uint32_t a_hi = tmem_d + BLOCK_N + s * 64 + k * UMMA_K, a_lo = a_hi + 32;
            mma_tf32_ts(tmem_d, a_lo, b_hi, idesc, (it | k) != 0);
            mma_tf32_ts(tmem_d, a_hi, b_lo, idesc, 1);
            mma_tf32_ts(tmem_d, a_hi, b_hi, idesc, 1);
          }
        }
        mma_commit(empty(s));
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(tmem_full);
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;  // column n0 + t of g  <->  TMEM lane t   (warps 2-5)
    const int tt = threadIdx.x - 64;    // 0..255 over all eight transform warps
    const bool a_warp = warp < 6;       // these own a TMEM lane quarter: g^T operand and epilogue
    const bool a_live = a_warp && quarter * 32 < p.a_cols;  // warp-uniform: this quarter holds real columns of g
    float bias_acc = 0.f;
    if (a_warp && !a_live) {  // lanes beyond the loaded columns: zero operand rows, written once for every stage
      float z[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) z[r] = 0.f;
      for (int c = 0; c < STAGES * S::A_COLS; c += 32)
        tmem_st_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + BLOCK_N + c, z);
      tmem_st_wait();
    }
    const float sg = F16 ? pow2f(f16_scale_exp(*p.g_absmax)) : 1.f;
    const float sx = F16 ? pow2f(f16_scale_exp(*p.x_absmax)) : 1.f;
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(full_tma(s), ph);
      if (F16) {
        if (a_live) {
          // A: column t of the [32 rows][a_cols] tile -> TMEM lane t as 16 + 16 packed fp16 pairs (rows 2c, 2c+1)
          const float* gcol = reinterpret_cast<const float*>(base_ptr + s * S::STAGE_BYTES) + t;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int r = 0; r < WG_ROWS; r += 2) {
            const float v0 = gcol[r * p.a_cols], v1 = gcol[(r + 1) * p.a_cols];
            bias_acc += v0;
            bias_acc += v1;
            split_f16x2(v0 * sg, v1 * sg, hi[r >> 1], lo[r >> 1]);
          }
          const uint32_t a_dst = tmem_d + ((uint32_t)(quarter * 32) << 16) + BLOCK_N + s * S::A_COLS;
          tmem_st_32x16(a_dst, hi);
          tmem_st_32x16(a_dst + 16, lo);
        }
        // B: fp32 [4 boxes][32 rows][32 floats] -> fp16 hi / lo, MN-major SWIZZLE_128B
        const float4* raw = reinterpret_cast<const float4*>(base_ptr + s * S::STAGE_BYTES + WG_A_BYTES);
        uint8_t* bhi = base_ptr + s * S::STAGE_BYTES + WG_A_BYTES + B_B;
        uint8_t* blo = bhi + S::B16_B;
#pragma unroll
        for (int i = 0; i < B_B / 16 / 256; ++i) {
          // 16 consecutive lanes take one row of a box PAIR (2 x 8 float4 in, one whole 128-byte fp16 row out):
          // quarter-warps read 128 contiguous bytes, half-warps write 128 contiguous (swizzled) bytes
          const int unit = (tt >> 4) + 16 * i, r = unit & 31, j = tt & 15;
          const int box = 2 * (unit >> 5) + (j >> 3), q = j & 7;
          float4 v = raw[box * 256 + r * 8 + q];
          if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          uint2 h, l;
          split_f16x2(v.x * sx, v.y * sx, h.x, l.x);
          split_f16x2(v.z * sx, v.w * sx, h.y, l.y);
          const int chunk = ((box & 1) << 2) | (q >> 1);  // 16-byte chunk of the 128-byte row (64 fp16 along k_in)
          const int off = (box >> 1) * 4096 + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + ((q & 1) << 3);
          *reinterpret_cast<uint2*>(bhi + off) = h;
          *reinterpret_cast<uint2*>(blo + off) = l;
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(full_ab(s));
        continue;
      }
      if (a_live) {
        // A: column t of the [32 rows][a_cols] tile -> TMEM lane t (hi | lo), plus the bias gradient
        const float* gcol = reinterpret_cast<const float*>(base_ptr + s * S::STAGE_BYTES) + t;
        float hi[32], lo[32];
#pragma unroll
        for (int r = 0; r < WG_ROWS; ++r) {
          const float v = gcol[r * p.a_cols];
          bias_acc += v;
          split_tf32(v, hi[r], lo[r]);
        }
        const uint32_t a_dst = tmem_d + ((uint32_t)(quarter * 32) << 16) + BLOCK_N + s * 64;
        tmem_st_32x32(a_dst, hi);
        tmem_st_32x32(a_dst + 32, lo);
      }
      // B: element-wise split in place (independent of the swizzled placement)
      float4* bhi = reinterpret_cast<float4*>(base_ptr + s * S::STAGE_BYTES + WG_A_BYTES);
      float4* blo = reinterpret_cast<float4*>(base_ptr + s * S::STAGE_BYTES + WG_A_BYTES + B_B);
#pragma unroll
      for (int c = tt; c < B_B / 16; c += 256) {
        float4 v = bhi[c];
        if (p.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        bhi[c] = h;
        blo[c] = l;
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(full_ab(s));
    }
    const int n = n0 + t;
    if (a_warp && p.partial_bias && blockIdx.y == 0 && n < p.n_out) p.partial_bias[(int64_t)blockIdx.z * p.n_out + n] = bias_acc;
    float* dst = p.partial + ((int64_t)blockIdx.z * p.n_out + n) * p.k_in;
    if (a_warp && n_iter > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; a_warp && c0 < BLOCK_N; c0 += 32) {
      float v[32];
      if (n_iter > 0) {
        __syncwarp();
        tmem_ld_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (F16) {
        const float inv = pow2f(-(f16_scale_exp(*p.g_absmax) + f16_scale_exp(*p.x_absmax)));
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= inv;
      }
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        const int k = k0 + c0 + gq * 4;
        if (n < p.n_out && k < p.k_in) st4(dst + k, make_float4(v[gq * 4], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, S::TMEM_COLS);
}

// ---- weight gradient over CTA PAIRS (fp16x3, n_out and k_in multiples of 256) --------------------------
// wgrad_x3_kernel<128, true> converts 8192 operand elements (g 32 x 128 and x 32 x 128) for every 128 x 128 x 32
// MMA triple: ~2100 warp instructions against the 4 x 384 issue slots that the MMAs of a stage last -- it is
// issue-bound by construction (profiles/r02_ncu_gemm_full.txt: tensor pipe 37 % active at 58 % issue utilisation).
// Here the two CTAs of a cluster (the two SMs of a TPC) own one 256 x 256 tile of dW for a slice of the rows:
//   A = g^T, M = 256: each CTA splits ITS 128 columns of g into its own tensor memory (lanes = rows of dW),
//   B = x,   N = 256: each CTA converts ITS 128 columns of x into its own shared memory (tcgen05.mma.cta_group::2
//                     reads half of the N extent from either CTA),
//   D: 128 lanes x 256 fp32 columns in either CTA.
// The same 8192 conversions per CTA and stage now feed 128 x 256 x 32 x 3 of MMA work per SM (768 cycles): half the
// conversion work and half the L2 -> SM bytes per MMA cycle, and four 48 KB stages fit beside the accumulator.
// Warps: 0 producer (own tiles, own barrier), 1 MMA issuer (leader CTA only), 2-5 g^T -> tensor memory + bias
// gradient, 6-9 x -> fp16 hi / lo MN-major tiles; all eight drain the accumulator (32-column blocks 0-3 / 4-7).
#ifndef T2H_WGP_STAGES
#define T2H_WGP_STAGES 6   // measured: 3 -> 4 stages 3-13 % faster (the life of a stage is a memory latency + conversion + MMAs)
#endif
constexpr int WGP_STAGES = T2H_WGP_STAGES;
constexpr int WGP_X_RAW = 4 * WG_GROUP_BYTES;            // this CTA's 128 columns of x: four 32-float boxes
constexpr int WGP_X16 = WG_ROWS * 128 * 2;               // one fp16 operand tile
// g raw | x: the fp32 tile is converted IN PLACE to (x hi | x lo) -- same 16 KB; the four converting warps hold the
// whole tile in registers across a named barrier before the first of them writes
constexpr int WGP_STAGE_BYTES = WG_A_BYTES + WGP_X_RAW;
static_assert(2 * WGP_X16 == WGP_X_RAW, "in-place conversion: fp16 hi + lo fill the fp32 tile");
constexpr int WGP_TOTAL = WGP_STAGES * WGP_STAGE_BYTES + 256 + 1024;
constexpr int WGP_TILE = 256;                            // dW tile edge (both ways)
constexpr int WGP_A_COLS = 32;                           // TMEM columns of (g hi | g lo) per stage
constexpr int WGP_TMEM_COLS = 512;                       // 256 accumulator + stages x 32 operand columns
static_assert(WGP_TILE + WGP_STAGES * WGP_A_COLS <= WGP_TMEM_COLS, "tensor memory");
static_assert(WGP_TOTAL <= 227 * 1024, "shared memory");

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_f16_pair_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x, const WgradArgs p,
                      const int tiles_k, const int n_tiles) {
  constexpr int STAGES = WGP_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * WGP_STAGE_BYTES;
  auto full_tma = [&](int s) { return bars + 8u * s; };                  // this CTA's tiles have landed
  auto full_ab = [&](int s) { return bars + 8u * (STAGES + s); };        // both CTAs' operands are converted (leader's)
  auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };      // the pair's MMAs have read the stage
  const uint32_t tmem_full = bars + 8u * (3 * STAGES);
  const uint32_t tmem_slot = bars + kTmemSlotOffset;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * WGP_STAGE_BYTES + kTmemSlotOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1);
  const int tile = pair % n_tiles, split = pair / n_tiles;
  const int n0 = (tile / tiles_k) * WGP_TILE + (int)rank * 128;   // this CTA's columns of g = its rows of dW
  const int kd0 = (tile % tiles_k) * WGP_TILE;                    // columns of dW held by the accumulator
  const int kx0 = kd0 + (int)rank * 128;                          // this CTA's columns of x
  const int64_t r_begin = (int64_t)split * p.rows_per_split;
  const int64_t r_end = min(r_begin + p.rows_per_split, p.rows);
  const int n_iter = r_end > r_begin ? (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_tma(s), 1);
      mbar_init(full_ab(s), 16);   // one arrival per transform warp of both CTAs
      mbar_init(empty(s), 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, WGP_TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(empty(s), ph ^ 1);
      if (elect_one()) {
        const uint32_t stage = base + s * WGP_STAGE_BYTES;
        const int r = (int)(r_begin + (int64_t)it * WG_ROWS);
        mbar_arrive_expect_tx(full_tma(s), (uint32_t)(WG_A_BYTES + WGP_X_RAW));
        tma_load_2d(stage, &tm_g, full_tma(s), n0, r);   // the global tail is zero-filled by TMA
#pragma unroll
        for (int gq = 0; gq < 4; ++gq)
          tma_load_2d(stage + WG_A_BYTES + gq * WG_GROUP_BYTES, &tm_x, full_tma(s), kx0 + gq * 32, r);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (rank == 0) {  // the leader's MMAs drive the tensor cores of both SMs
      constexpr uint32_t idesc = make_idesc_f16(2 * BLOCK_M, WGP_TILE, 0, 1);  // A from TMEM, B MN-major
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait_cluster(full_ab(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi0 = base + s * WGP_STAGE_BYTES + WG_A_BYTES;
          const uint64_t d_hi0 = make_smem_desc(b_hi0, 4096, 1024, kLayoutSW128);
          const uint64_t d_lo0 = make_smem_desc(b_hi0 + WGP_X16, 4096, 1024, kLayoutSW128);
#pragma unroll
          for (int k = 0; k < WG_ROWS / 16; ++k) {
            const uint64_t b_hi = d_hi0 + (uint64_t)(k * 128), b_lo = d_lo0 + (uint64_t)(k * 128);  // two 8-row atoms
            const uint32_t a_hi = tmem_d + WGP_TILE + s * WGP_A_COLS + k * 8, a_lo = a_hi + 16;
            mma_f16_ts_pair(tmem_d, a_lo, b_hi, idesc, (it | k) != 0);
            mma_f16_ts_pair(tmem_d, a_hi, b_lo, idesc, 1);
            mma_f16_ts_pair(tmem_d, a_hi, b_hi, idesc, 1);
          }
          mma_commit_pair(empty(s));
        }
        __syncwarp();
      }
      if (n_iter > 0 && elect_one()) mma_commit_pair(tmem_full);
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;  // TMEM lane = column n0 + t of g
    const bool a_warp = warp < 6;
    const float sg = pow2f(f16_scale_exp(*p.g_absmax));
    const float sx = pow2f(f16_scale_exp(*p.x_absmax));
    float bias_acc = 0.f;
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(full_tma(s), ph);
      if (a_warp) {
        // column t of the [32 rows][128] tile -> TMEM lane t as 16 + 16 packed fp16 pairs (rows 2c, 2c + 1)
        const float* gcol = reinterpret_cast<const float*>(base_ptr + s * WGP_STAGE_BYTES) + t;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int r = 0; r < WG_ROWS; r += 2) {
          const float v0 = gcol[r * 128], v1 = gcol[(r + 1) * 128];
          bias_acc += v0;
          bias_acc += v1;
          split_f16x2(v0 * sg, v1 * sg, hi[r >> 1], lo[r >> 1]);
        }
        const uint32_t a_dst = tmem_d + ((uint32_t)(quarter * 32) << 16) + WGP_TILE + s * WGP_A_COLS;
        tmem_st_32x16(a_dst, hi);
        tmem_st_32x16(a_dst + 16, lo);
        tmem_st_wait();
        tc_fence_before();
      } else {
        // fp32 [4 boxes][32 rows][32 floats] -> fp16 hi / lo, MN-major SWIZZLE_128B (two 64-wide groups 4096 B apart);
        // 16 consecutive lanes take one row of a box pair: 128 contiguous bytes in, one swizzled 128-byte row out
        const int tb = threadIdx.x - 192;
        const float4* raw = reinterpret_cast<const float4*>(base_ptr + s * WGP_STAGE_BYTES + WG_A_BYTES);
        uint8_t* bhi = base_ptr + s * WGP_STAGE_BYTES + WG_A_BYTES;
        uint8_t* blo = bhi + WGP_X16;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int unit = (tb >> 4) + 8 * i, r = unit & 31, j = tb & 15;
          const int box = 2 * (unit >> 5) + (j >> 3), q = j & 7;
          v[i] = raw[box * 256 + r * 8 + q];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four x warps: the fp32 tile is in registers
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int unit = (tb >> 4) + 8 * i, r = unit & 31, j = tb & 15;
          const int box = 2 * (unit >> 5) + (j >> 3), q = j & 7;
          if (p.relu_in) { v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f); v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f); }
          uint2 h, l;
          split_f16x2(v[i].x * sx, v[i].y * sx, h.x, l.x);
          split_f16x2(v[i].z * sx, v[i].w * sx, h.y, l.y);
          const int chunk = ((box & 1) << 2) | (q >> 1);  // 16-byte chunk of the 128-byte row (64 fp16 along k_in)
          const int off = (box >> 1) * 4096 + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + ((q & 1) << 3);
          *reinterpret_cast<uint2*>(bhi + off) = h;
          *reinterpret_cast<uint2*>(blo + off) = l;
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) {  // one (possibly remote) arrival per warp
        if (rank != 0) mbar_arrive_remote(full_ab(s), 0);
        else mbar_arrive(full_ab(s));
      }
    }
    const int n = n0 + t;
    if (a_warp && p.partial_bias && kd0 == 0) p.partial_bias[(int64_t)split * p.n_out + n] = bias_acc;
    float* dst = p.partial + ((int64_t)split * p.n_out + n) * p.k_in + kd0;
    if (n_iter > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    const float inv = pow2f(-(f16_scale_exp(*p.g_absmax) + f16_scale_exp(*p.x_absmax)));
#pragma unroll 1
    for (int c0 = a_warp ? 0 : 128; c0 < (a_warp ? 128 : 256); c0 += 32) {
      float v[32];
      if (n_iter > 0) {
        __syncwarp();
        tmem_ld_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int gq = 0; gq < 8; ++gq)
        st4(dst + c0 + gq * 4, make_float4(v[gq * 4] * inv, v[gq * 4 + 1] * inv, v[gq * 4 + 2] * inv, v[gq * 4 + 3] * inv));
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's shared and tensor memory are not read by the pair's MMAs any more
  if (warp == 1) tmem_dealloc_pair(tmem_d, WGP_TMEM_COLS);
}

// dst[n, k] (ld) = sum over splits of partial[s, n, k]; TPO threads share one float4 of the output
// (narrow layers have few outputs but hundreds of row splits), fixed strided order + xor tree
template <int TPO>
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias,
                                    int splits, int n_out, int k_in, float* __restrict__ dst, int64_t ld,
                                    float* __restrict__ dst_bias) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = gid / TPO;  // float4 index
  const int part = (int)(gid % TPO);
  const int64_t total4 = (int64_t)n_out * k_in / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float b = 0.f;
  const bool has_w = i < total4, has_b = dst_bias && i < n_out;
  for (int s = part; s < splits; s += TPO) {
    if (has_w) {
      const float4 v = ld4(partial + ((int64_t)s * n_out * k_in) + i * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (has_b) b += partial_bias[(int64_t)s * n_out + i];
  }
#pragma unroll
  for (int off = 1; off < TPO; off <<= 1) {
    const float4 o = shfl_xor4(acc, off);
    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  if (part != 0) return;
  if (has_b) dst_bias[i] = b;
  if (has_w) {
    const int64_t e = i * 4;
    st4(dst + (e / k_in) * ld + (e % k_in), acc);
  }
}

static void launch_wgrad_reduce(const float* partial, const float* partial_bias, int splits, int n_out, int k_in,
                                float* dst, int64_t ld, float* dst_bias, cudaStream_t s) {
  const int64_t total4 = (int64_t)n_out * k_in / 4;
  const int64_t outs = total4 > n_out ? total4 : n_out;
  if (splits >= 64) {
    wgrad_reduce_kernel<32><<<(unsigned)((outs * 32 + 255) / 256), 256, 0, s>>>(partial, partial_bias, splits, n_out, k_in, dst, ld, dst_bias);
  } else if (splits >= 16) {
    wgrad_reduce_kernel<8><<<(unsigned)((outs * 8 + 255) / 256), 256, 0, s>>>(partial, partial_bias, splits, n_out, k_in, dst, ld, dst_bias);
  } else {
    wgrad_reduce_kernel<1><<<(unsigned)((outs + 255) / 256), 256, 0, s>>>(partial, partial_bias, splits, n_out, k_in, dst, ld, dst_bias);
  }
}

// column sums (bias gradient): partial[b, c] = sum of rows [b*rpb, (b+1)*rpb) of g[:, c]
template <int CW>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ g, int64_t rows, int n, int64_t ld,
                                                             int64_t rows_per_block, float* __restrict__ partial) {
  __shared__ float red[256];
  constexpr int RY = 256 / CW;
  const int tx = threadIdx.x % CW, ty = threadIdx.x / CW;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
  for (int c = tx + blockIdx.y * CW; c < n; c += CW * gridDim.y) {
    float acc = 0.f;
    for (int64_t r = r0 + ty; r < r1; r += RY) acc += g[r * ld + c];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (ty == 0) {
      for (int j = 1; j < RY; ++j) acc += red[j * CW + tx];
      partial[(int64_t)blockIdx.x * n + c] = acc;
    }
    __syncthreads();
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int blocks, int n, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float acc = 0.f;
  for (int b = 0; b < blocks; ++b) acc += partial[(int64_t)b * n + c];
  out[c] = acc;
}

__global__ void split_tf32_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float h, l;
  split_tf32(w[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

// max |x| over a (rows, k) matrix with row pitch ld, as the bit pattern of the (non-negative) float, merged into
// *slot with atomicMax (unsigned order == float order for non-negative values; NaN sorts above infinity)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, int64_t ld, int k4, int64_t n4, uint32_t* __restrict__ slot) {
  float m = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool flat = (ld == (int64_t)k4 * 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      if (j < n4) {
        const int64_t off = flat ? j * 4 : (j / k4) * ld + (j % k4) * 4;
        v[u] = __ldg(reinterpret_cast<const float4*>(x + off));
      } else v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // integer max of the magnitudes' bit patterns keeps NaN visible (fmaxf would drop it)
      m = __uint_as_float(max(max(__float_as_uint(fabsf(v[u].x)), __float_as_uint(fabsf(v[u].y))),
                              max(max(__float_as_uint(fabsf(v[u].z)), __float_as_uint(fabsf(v[u].w))), __float_as_uint(m))));
    }
  }
  uint32_t b = __float_as_uint(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  __shared__ uint32_t warp_max[8];
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x < 8) {
    b = warp_max[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffu, b, o));
    if (threadIdx.x == 0 && b) atomicMax(slot, b);
  }
}

// out = a + b with max |out| merged into *slot: the sum of two gradient branches is the operand of the next
// input-gradient GEMM, so its maximum is taken while it is being written instead of in a second pass
__global__ void __launch_bounds__(256) add_absmax_kernel(const float4* __restrict__ a, const float4* __restrict__ b, int64_t n4,
                                                         float4* __restrict__ out, uint32_t* __restrict__ slot) {
  float m = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 4 * stride) {
    float4 va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      if (j < n4) { va[u] = __ldg(a + j); vb[u] = __ldg(b + j); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = i + u * stride;
      if (j < n4) {
        const float4 o = make_float4(va[u].x + vb[u].x, va[u].y + vb[u].y, va[u].z + vb[u].z, va[u].w + vb[u].w);
        out[j] = o;
        m = __uint_as_float(max(max(__float_as_uint(fabsf(o.x)), __float_as_uint(fabsf(o.y))),
                                max(max(__float_as_uint(fabsf(o.z)), __float_as_uint(fabsf(o.w))), __float_as_uint(m))));
      }
    }
  }
  uint32_t bits = __float_as_uint(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  __shared__ uint32_t warp_max[8];
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x < 8) {
    bits = warp_max[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffu, bits, o));
    if (threadIdx.x == 0 && bits) atomicMax(slot, bits);
  }
}

__global__ void split_f16_kernel(const float* __restrict__ w, int64_t n2, const uint32_t* __restrict__ absmax,
                                 uint32_t* __restrict__ hi, uint32_t* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const float sc = pow2f(f16_scale_exp(*absmax));
  const float2 v = reinterpret_cast<const float2*>(w)[i];
  uint32_t h, l;
  split_f16x2(v.x * sc, v.y * sc, h, l);
  hi[i] = h;
  lo[i] = l;
}

// ---- host side ----------------------------------------------------------------------------------
// Kernel-variant switches of the design study.  The product build fixes them at compile time -- the
// entry points read no environment and keep no state; -DT2H_ABLATION_ENV re-enables the environment overrides.
static inline int ablation_switch(const char* name, int product_value) {
#ifdef T2H_ABLATION_ENV
  const char* e = getenv(name);
  return e ? atoi(e) : product_value;
#else
  (void)name;
  return product_value;
#endif
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver entry point is looked up once (an immutable function pointer, initialised thread-safely)
static EncodeTiledFn encode_fn() {
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? (EncodeTiledFn)ptr : nullptr;
  }();
  return fn;
}

// 2-D fp32 tensor [outer, inner] (inner contiguous), box [box_outer, box_inner], 128-byte swizzle
static bool make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                     uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D fp16 weight matrix [outer, inner] (K contiguous), box [box_outer, 64] = 128-byte swizzle rows
static bool make_map_f16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// channels-last plane (B, H, W, C) as a 4-D tensor (C, W, H, B); box = 32 channels x bw x bh pixels of one image
static bool make_map_4d(CUtensorMap* map, const float* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t B,
                        uint32_t box_c, uint32_t bw, uint32_t bh, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * sizeof(float), W * C * sizeof(float), H * W * C * sizeof(float)};
  cuuint32_t box[4] = {box_c, bw, bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PlaneGeom { int B, H, W; };  // conv mode only

template <bool F16, int BLOCK_N, int WSTEPS = 0, bool PAIR = false, bool DS = false>
static int launch_linear_persistent(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& whi, const CUtensorMap& wlo,
                                    const CUtensorMap& mout, const CUtensorMap& maux, const LinearArgs& args,
                                    cudaStream_t stream) {
  auto kern = linear_x3_persistent_kernel<F16, BLOCK_N, WSTEPS, PAIR, DS>;
  using S = PSmem<F16, BLOCK_N, WSTEPS, PAIR, DS>;
  // per device and idempotent; set on every launch so that the entry point keeps no state
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  const int64_t tiles = ((args.rows + BLOCK_M - 1) / BLOCK_M) * ((args.n_out + BLOCK_N - 1) / BLOCK_N);
  unsigned grid = (unsigned)(tiles < kSMs ? tiles : kSMs);
  if (WSTEPS > 0) {  // a multiple of the N-tile count: `tile += gridDim.x` then keeps each CTA on one N-tile
    const unsigned n_tiles = (unsigned)((args.n_out + BLOCK_N - 1) / BLOCK_N);
    grid = (grid / n_tiles) * n_tiles;
  }
  if (PAIR) {
    // pairs of CTAs (cluster of 2 = the two SMs of a TPC) walk the list of 256-row tiles
    const int64_t m_tiles = (args.rows + BLOCK_M - 1) / BLOCK_M, n_tiles = (args.n_out + BLOCK_N - 1) / BLOCK_N;
    const int64_t pair_tiles = ((m_tiles + 1) / 2) * n_tiles;
    const unsigned pairs = (unsigned)(pair_tiles < kSMs / 2 ? pair_tiles : kSMs / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, x1, x2, whi, wlo, mout, maux, args, pair_tiles) != cudaSuccess) {
      (void)cudaGetLastError();
      return T2H_ERR_CUDA;
    }
    return T2H_OK;
  }
  kern<<<grid, P_THREADS, S::TOTAL, stream>>>(x1, x2, whi, wlo, mout, maux, args, tiles);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

template <int BLOCK_N, bool A_TMEM, int STAGES_>
static int launch_linear(const CUtensorMap& x1, const CUtensorMap& x2, const float* w_hi, const float* w_lo, int k_total,
                         const LinearArgs& args, cudaStream_t stream, const PlaneGeom* pg = nullptr) {
  CUtensorMap whi, wlo, mout, maux;
  if (!make_map(&whi, w_hi, k_total, args.n_out, k_total, BLOCK_K, BLOCK_N)) return T2H_ERR_CUDA;
  if (!make_map(&wlo, w_lo, k_total, args.n_out, k_total, BLOCK_K, BLOCK_N)) return T2H_ERR_CUDA;
  if (pg) {
    if (!make_map_4d(&mout, args.out, args.n_out, pg->W, pg->H, pg->B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
    maux = mout;
    if (args.aux_kind == 1 && !make_map_4d(&maux, args.mask, args.n_out, pg->W, pg->H, pg->B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
    if (args.aux_kind == 2 && !make_map_4d(&maux, args.residual, args.n_out, pg->W, pg->H, pg->B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  } else {
    if (!make_map(&mout, args.out, args.n_out, args.rows, args.ld_out, 32, BLOCK_M)) return T2H_ERR_CUDA;
    maux = mout;
    if (args.aux_kind == 1 && !make_map(&maux, args.mask, args.n_out, args.rows, args.ld_mask, 32, BLOCK_M)) return T2H_ERR_CUDA;
    if (args.aux_kind == 2 && !make_map(&maux, args.residual, args.n_out, args.rows, args.ld_res, 32, BLOCK_M)) return T2H_ERR_CUDA;
  }
  {
    // T2H_LINEAR_NONPERSISTENT=1 (ablation): the one-tile-per-CTA kernels below; =2: only for the narrow tiles
    const int one_tile = ablation_switch("T2H_LINEAR_NONPERSISTENT", 0);
    if (one_tile == 0 || (one_tile == 2 && BLOCK_N == 128))
      return launch_linear_persistent<false, BLOCK_N>(x1, x2, whi, wlo, mout, maux, args, stream);
  }
  auto kern = linear_tf32x3_kernel<BLOCK_N, A_TMEM, STAGES_>;
  // per device and idempotent; set on every launch so that the entry point keeps no state
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BLOCK_N, A_TMEM, STAGES_>::TOTAL) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  const int64_t tiles = ((args.rows + BLOCK_M - 1) / BLOCK_M) * ((args.n_out + BLOCK_N - 1) / BLOCK_N);
  if (tiles > 0x7fffffffLL) return T2H_ERR_UNSUPPORTED_SHAPE;
  kern<<<(unsigned)tiles, kThreads, Smem<BLOCK_N, A_TMEM, STAGES_>::TOTAL, stream>>>(x1, x2, whi, wlo, mout, maux, args);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

}  // namespace gemm
}  // namespace t2h

using namespace t2h;
using namespace t2h::gemm;

extern "C" int t2h_split_tf32(const float* w, int64_t n, float* hi, float* lo, t2h_stream_t stream) {
  if (!w || !hi || !lo || n < 0) return T2H_ERR_INVALID_ARGUMENT;
  if (n == 0) return T2H_OK;
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, n, hi, lo);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_linear_fwd(const float* x1, int64_t ld_x1, int k1, const float* x2, int64_t ld_x2, int k2,
                              int64_t rows, const float* w_hi, const float* w_lo, int n_out, const float* bias,
                              int relu_in, const float* mask, int64_t ld_mask, const float* residual, int64_t ld_res,
                              float* out, int64_t ld_out, t2h_stream_t stream) {
  if (!x1 || !w_hi || !w_lo || !out || rows < 0 || k1 <= 0 || k2 < 0 || n_out <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (k2 > 0 && !x2) return T2H_ERR_INVALID_ARGUMENT;
  // TMA: 16-byte aligned bases and row pitches; a second source must start on a chunk boundary
  if ((k1 % 4) || (k2 % 4) || (ld_x1 % 4) || (k2 > 0 && ((ld_x2 % 4) || (k1 % BLOCK_K))) || (n_out % 4) ||
      (ld_out % 4) || (residual && (ld_res % 4)) || (mask && (ld_mask % 4)))
    return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out | (uintptr_t)bias |
       (uintptr_t)mask | (uintptr_t)residual) & 15)
    return T2H_ERR_INVALID_ARGUMENT;
  if (rows == 0) return T2H_OK;
  CUtensorMap m1, m2;
  if (!make_map(&m1, x1, k1, rows, ld_x1, BLOCK_K, BLOCK_M)) return T2H_ERR_CUDA;
  if (k2 > 0) { if (!make_map(&m2, x2, k2, rows, ld_x2, BLOCK_K, BLOCK_M)) return T2H_ERR_CUDA; }
  else m2 = m1;
  LinearArgs a;
  a.rows = rows; a.n_out = n_out;
  a.k1_chunks = (k1 + BLOCK_K - 1) / BLOCK_K;
  a.k_chunks = a.k1_chunks + (k2 + BLOCK_K - 1) / BLOCK_K;
  a.relu_in = relu_in; a.bias = bias; a.mask = mask; a.ld_mask = ld_mask;
  a.residual = residual; a.ld_res = ld_res; a.out = out; a.ld_out = ld_out;
  a.aux_kind = mask ? 1 : (residual ? 2 : 0);
  a.conv = 0; a.tiles_x = a.tiles_per_img = a.cin_chunks = 0;
  a.x_absmax = a.w_absmax = nullptr; a.out_absmax = nullptr;
  const int k_total = k1 + k2;
  cudaStream_t s = (cudaStream_t)stream;
  const int ss_only = ablation_switch("T2H_LINEAR_SS", 0), deep = ablation_switch("T2H_LINEAR_DEEP", 0);
  const bool shallow = a.k_chunks <= 2 && !deep;
  if (n_out <= 32) return shallow ? launch_linear<32, false, 1>(m1, m2, w_hi, w_lo, k_total, a, s)
                                  : launch_linear<32, false, 2>(m1, m2, w_hi, w_lo, k_total, a, s);
  if (n_out <= 64) return shallow ? launch_linear<64, false, 1>(m1, m2, w_hi, w_lo, k_total, a, s)
                                  : launch_linear<64, false, 2>(m1, m2, w_hi, w_lo, k_total, a, s);
  if (ss_only) return launch_linear<128, false, 3>(m1, m2, w_hi, w_lo, k_total, a, s);
  return launch_linear<128, true, 2>(m1, m2, w_hi, w_lo, k_total, a, s);
}


extern "C" int t2h_absmax(const float* x1, int64_t ld_x1, int k1, const float* x2, int64_t ld_x2, int k2, int64_t rows,
                          uint32_t* slot, t2h_stream_t stream) {
  if (!x1 || !slot || rows < 0 || k1 <= 0 || k2 < 0 || (k2 > 0 && !x2)) return T2H_ERR_INVALID_ARGUMENT;
  if ((k1 % 4) || (k2 % 4) || (ld_x1 % 4) || (k2 > 0 && (ld_x2 % 4))) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)x1 | (uintptr_t)x2) & 15) return T2H_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(slot, 0, sizeof(uint32_t), s) != cudaSuccess) { (void)cudaGetLastError(); return T2H_ERR_CUDA; }
  const float* xs[2] = {x1, x2};
  const int64_t lds[2] = {ld_x1, ld_x2};
  const int ks[2] = {k1, k2};
  for (int i = 0; i < 2; ++i) {
    if (ks[i] == 0 || rows == 0) continue;
    const int64_t n4 = rows * (ks[i] / 4);
    int64_t blocks = (n4 + 1023) / 1024;  // four float4 per thread and sweep
    if (blocks > 8 * kSMs) blocks = 8 * kSMs;
    absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(xs[i], lds[i], ks[i] / 4, n4, slot);
    T2H_CHECK_LAUNCH();
  }
  return T2H_OK;
}

extern "C" int t2h_add_absmax(const float* a, const float* b, int64_t n, float* out, uint32_t* slot, t2h_stream_t stream) {
  if (!a || !b || !out || !slot || n < 0) return T2H_ERR_INVALID_ARGUMENT;
  if (n % 4) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) return T2H_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemsetAsync(slot, 0, sizeof(uint32_t), s) != cudaSuccess) { (void)cudaGetLastError(); return T2H_ERR_CUDA; }
  if (n == 0) return T2H_OK;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 1023) / 1024;
  if (blocks > 8 * kSMs) blocks = 8 * kSMs;
  add_absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), n4,
                                                     reinterpret_cast<float4*>(out), slot);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_split_f16(const float* w, int64_t n, const uint32_t* absmax, uint16_t* hi, uint16_t* lo, t2h_stream_t stream) {
  if (!w || !absmax || !hi || !lo || n < 0) return T2H_ERR_INVALID_ARGUMENT;
  if (n % 2) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)w & 7) || (((uintptr_t)hi | (uintptr_t)lo) & 3)) return T2H_ERR_INVALID_ARGUMENT;
  if (n == 0) return T2H_OK;
  split_f16_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, n / 2, absmax, reinterpret_cast<uint32_t*>(hi), reinterpret_cast<uint32_t*>(lo));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_linear_fwd_f16(const float* x1, int64_t ld_x1, int k1, const float* x2, int64_t ld_x2, int k2,
                                  int64_t rows, const uint32_t* x_absmax, const uint16_t* w_hi, const uint16_t* w_lo,
                                  const uint32_t* w_absmax, int n_out, const float* bias, int relu_in, const float* mask,
                                  int64_t ld_mask, const float* residual, int64_t ld_res, float* out, int64_t ld_out,
                                  uint32_t* out_absmax, t2h_stream_t stream) {
  if (!x1 || !w_hi || !w_lo || !x_absmax || !w_absmax || !out || rows < 0 || k1 <= 0 || k2 < 0 || n_out <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if (k2 > 0 && !x2) return T2H_ERR_INVALID_ARGUMENT;
  // TMA: 16-byte aligned bases and row pitches (fp16 weight rows: K % 8); a second source starts on a chunk boundary
  if ((k1 % 4) || (k2 % 4) || ((k1 + k2) % 8) || (ld_x1 % 4) || (k2 > 0 && ((ld_x2 % 4) || (k1 % BLOCK_K))) || (n_out % 4) ||
      (ld_out % 4) || (residual && (ld_res % 4)) || (mask && (ld_mask % 4)))
    return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out | (uintptr_t)bias |
       (uintptr_t)mask | (uintptr_t)residual) & 15)
    return T2H_ERR_INVALID_ARGUMENT;
  if (rows == 0) return T2H_OK;
  CUtensorMap m1, m2, whi, wlo, mout, maux;
  if (!make_map(&m1, x1, k1, rows, ld_x1, BLOCK_K, BLOCK_M)) return T2H_ERR_CUDA;
  if (k2 > 0) { if (!make_map(&m2, x2, k2, rows, ld_x2, BLOCK_K, BLOCK_M)) return T2H_ERR_CUDA; }
  else m2 = m1;
  const int k_total = k1 + k2;
  const int bn = n_out <= 32 ? 32 : (n_out <= 64 ? 64 : 128);
  if (!make_map_f16(&whi, w_hi, k_total, n_out, bn) || !make_map_f16(&wlo, w_lo, k_total, n_out, bn)) return T2H_ERR_CUDA;
  LinearArgs a;
  a.rows = rows; a.n_out = n_out;
  a.k1_chunks = (k1 + BLOCK_K - 1) / BLOCK_K;
  a.k_chunks = a.k1_chunks + (k2 + BLOCK_K - 1) / BLOCK_K;
  a.relu_in = relu_in; a.bias = bias; a.mask = mask; a.ld_mask = ld_mask;
  a.residual = residual; a.ld_res = ld_res; a.out = out; a.ld_out = ld_out;
  a.aux_kind = mask ? 1 : (residual ? 2 : 0);
  a.conv = 0; a.tiles_x = a.tiles_per_img = a.cin_chunks = 0;
  a.x_absmax = x_absmax; a.w_absmax = w_absmax; a.out_absmax = out_absmax;
  if (out_absmax && cudaMemsetAsync(out_absmax, 0, sizeof(uint32_t), (cudaStream_t)stream) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  if (!make_map(&mout, out, n_out, rows, ld_out, 32, BLOCK_M)) return T2H_ERR_CUDA;
  maux = mout;
  if (a.aux_kind == 1 && !make_map(&maux, mask, n_out, rows, ld_mask, 32, BLOCK_M)) return T2H_ERR_CUDA;
  if (a.aux_kind == 2 && !make_map(&maux, residual, n_out, rows, ld_res, 32, BLOCK_M)) return T2H_ERR_CUDA;
  if (bn == 32) return launch_linear_persistent<true, 32>(m1, m2, whi, wlo, mout, maux, a, (cudaStream_t)stream);
  if (bn == 64) return launch_linear_persistent<true, 64>(m1, m2, whi, wlo, mout, maux, a, (cudaStream_t)stream);
  // CTA pairs (cta_group::2) for 128-wide tiles; T2H_LINEAR_PAIR=0 falls back to one CTA per tile (ablation).
  // The weight maps then have 64-row boxes (each CTA of a pair stages half of the weight tile).
  const int pair = ablation_switch("T2H_LINEAR_PAIR", 1);
  if (pair && rows >= 2 * BLOCK_M) {
    CUtensorMap whi2, wlo2;
    // Tile and epilogue flavour (measured on 2^20 rows, profiles/r02_gemm_probe_tiles.txt):
    //   256 x 256 tiles per pair (each CTA stages 128 weight rows) where n_out is a multiple of 256 and K > 128 --
    //   one conversion and one L2 -> SM transfer of x per 256 output columns (K = 1024 -> N = 512: 2.70 -> 2.34 ms;
    //   K = 128 -> N = 256 is bound by its epilogue and gets slower, 0.37 -> 0.43 ms);
    //   two staging slots per epilogue group where a mask / residual block is read per output block and a tile's
    //   MMAs are short (K <= 256: K = 256 -> N = 512 with a mask 1.26 -> 1.02 ms, K = 128 -> N = 256 0.73 -> 0.55 ms;
    //   with K >= 512 the deeper load ring is worth more than the second slot).
    int wide = n_out % 256 == 0 && a.k_chunks > 4;
    int ds = a.aux_kind != 0 && a.k_chunks <= 8;
    const int force_wide = ablation_switch("T2H_LINEAR_WIDE", -1), force_ds = ablation_switch("T2H_LINEAR_DS", -1);
    if (force_wide >= 0) wide = force_wide && n_out % 256 == 0;
    if (force_ds >= 0) ds = force_ds;
    if (wide) {
      if (!make_map_f16(&whi2, w_hi, k_total, n_out, 128) || !make_map_f16(&wlo2, w_lo, k_total, n_out, 128)) return T2H_ERR_CUDA;
      if (ds) return launch_linear_persistent<true, 256, 0, true, true>(m1, m2, whi2, wlo2, mout, maux, a, (cudaStream_t)stream);
      return launch_linear_persistent<true, 256, 0, true>(m1, m2, whi2, wlo2, mout, maux, a, (cudaStream_t)stream);
    }
    if (!make_map_f16(&whi2, w_hi, k_total, n_out, 64) || !make_map_f16(&wlo2, w_lo, k_total, n_out, 64)) return T2H_ERR_CUDA;
    if (ds) return launch_linear_persistent<true, 128, 0, true, true>(m1, m2, whi2, wlo2, mout, maux, a, (cudaStream_t)stream);
    return launch_linear_persistent<true, 128, 0, true>(m1, m2, whi2, wlo2, mout, maux, a, (cudaStream_t)stream);
  }
  // K <= 256: the weight tile of a CTA's N-tile can stay resident in shared memory (T2H_LINEAR_WRES=1).  OFF by
  // default -- measured: it halves the L2 -> SM tile traffic but leaves only 2 x-stages (64 KB in flight) for
  // K = 256, and the layer gets SLOWER (256 -> 512: 0.53 vs 0.44 ms), i.e. these layers are bound by the depth
  // of the load pipeline, not by the fabric; K = 128 (4 stages) is a wash (0.154 vs 0.157 ms).
  const int wres = ablation_switch("T2H_LINEAR_WRES", 0);
  const int n_steps = (a.k_chunks + 1) / 2, n_tiles128 = (n_out + 127) / 128;
  if (wres && n_tiles128 <= kSMs) {
    if (n_steps <= 2) return launch_linear_persistent<true, 128, 2>(m1, m2, whi, wlo, mout, maux, a, (cudaStream_t)stream);
    if (n_steps <= 4) return launch_linear_persistent<true, 128, 4>(m1, m2, whi, wlo, mout, maux, a, (cudaStream_t)stream);
  }
  return launch_linear_persistent<true, 128>(m1, m2, whi, wlo, mout, maux, a, (cudaStream_t)stream);
}

extern "C" int t2h_conv3x3_fwd(const float* x, int B, int H, int W, int cin, const float* w_hi, const float* w_lo,
                               int cout, const float* bias, int relu_in, const float* mask, const float* residual,
                               float* out, t2h_stream_t stream) {
  if (!x || !w_hi || !w_lo || !out || B < 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if ((cin % BLOCK_K) || (cout % 4) || (W % CONV_TW) || (H % CONV_TH)) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)x | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)mask |
       (uintptr_t)residual) & 15)
    return T2H_ERR_INVALID_ARGUMENT;
  if (B == 0) return T2H_OK;
  CUtensorMap mx;
  if (!make_map_4d(&mx, x, cin, W, H, B, BLOCK_K, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  LinearArgs a;
  a.rows = (int64_t)B * H * W; a.n_out = cout;
  a.cin_chunks = cin / BLOCK_K;
  a.k_chunks = a.k1_chunks = 9 * a.cin_chunks;
  a.relu_in = relu_in; a.bias = bias; a.mask = mask; a.ld_mask = cout; a.residual = residual; a.ld_res = cout;
  a.out = out; a.ld_out = cout;
  a.aux_kind = mask ? 1 : (residual ? 2 : 0);
  if (mask && residual) return T2H_ERR_UNSUPPORTED_SHAPE;  // only one epilogue operand is staged in conv mode
  a.conv = 1; a.tiles_x = W / CONV_TW; a.tiles_per_img = (H / CONV_TH) * a.tiles_x;
  a.x_absmax = a.w_absmax = nullptr; a.out_absmax = nullptr;
  PlaneGeom pg{B, H, W};
  cudaStream_t s = (cudaStream_t)stream;
  const int k_total = 9 * cin;
  if (cout <= 32) return launch_linear<32, false, 2>(mx, mx, w_hi, w_lo, k_total, a, s, &pg);
  if (cout <= 64) return launch_linear<64, false, 2>(mx, mx, w_hi, w_lo, k_total, a, s, &pg);
  return launch_linear<128, true, 2>(mx, mx, w_hi, w_lo, k_total, a, s, &pg);
}

extern "C" int t2h_conv3x3_fwd_f16(const float* x, int B, int H, int W, int cin, const uint32_t* x_absmax,
                                   const uint16_t* w_hi, const uint16_t* w_lo, const uint32_t* w_absmax, int cout,
                                   const float* bias, int relu_in, const float* mask, const float* residual, float* out,
                                   uint32_t* out_absmax, t2h_stream_t stream) {
  if (!x || !w_hi || !w_lo || !x_absmax || !w_absmax || !out || B < 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if ((cin % BLOCK_K) || (cout % 4) || (W % CONV_TW) || (H % CONV_TH)) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (mask && residual) return T2H_ERR_UNSUPPORTED_SHAPE;  // only one epilogue operand is staged in conv mode
  if (((uintptr_t)x | (uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)out | (uintptr_t)bias | (uintptr_t)mask |
       (uintptr_t)residual) & 15)
    return T2H_ERR_INVALID_ARGUMENT;
  if (B == 0) return T2H_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int k_total = 9 * cin;
  const int bn = cout <= 32 ? 32 : (cout <= 64 ? 64 : 128);
  CUtensorMap mx, whi, wlo, mout, maux;
  if (!make_map_4d(&mx, x, cin, W, H, B, BLOCK_K, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  if (!make_map_f16(&whi, w_hi, k_total, cout, bn) || !make_map_f16(&wlo, w_lo, k_total, cout, bn)) return T2H_ERR_CUDA;
  if (!make_map_4d(&mout, out, cout, W, H, B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  maux = mout;
  if (mask && !make_map_4d(&maux, mask, cout, W, H, B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  if (residual && !make_map_4d(&maux, residual, cout, W, H, B, 32, CONV_TW, CONV_TH)) return T2H_ERR_CUDA;
  LinearArgs a;
  a.rows = (int64_t)B * H * W; a.n_out = cout;
  a.cin_chunks = cin / BLOCK_K;
  a.k_chunks = a.k1_chunks = 9 * a.cin_chunks;
  a.relu_in = relu_in; a.bias = bias; a.mask = mask; a.ld_mask = cout; a.residual = residual; a.ld_res = cout;
  a.out = out; a.ld_out = cout;
  a.aux_kind = mask ? 1 : (residual ? 2 : 0);
  a.conv = 1; a.tiles_x = W / CONV_TW; a.tiles_per_img = (H / CONV_TH) * a.tiles_x;
  a.x_absmax = x_absmax; a.w_absmax = w_absmax; a.out_absmax = out_absmax;
  if (out_absmax && cudaMemsetAsync(out_absmax, 0, sizeof(uint32_t), s) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  if (bn == 32) return launch_linear_persistent<true, 32>(mx, mx, whi, wlo, mout, maux, a, s);
  if (bn == 64) return launch_linear_persistent<true, 64>(mx, mx, whi, wlo, mout, maux, a, s);
  const int pair = ablation_switch("T2H_LINEAR_PAIR", 1);
  if (pair && a.rows >= 2 * BLOCK_M) {
    CUtensorMap whi2, wlo2;
    if (!make_map_f16(&whi2, w_hi, k_total, cout, 64) || !make_map_f16(&wlo2, w_lo, k_total, cout, 64)) return T2H_ERR_CUDA;
    return launch_linear_persistent<true, 128, 0, true>(mx, mx, whi2, wlo2, mout, maux, a, s);
  }
  return launch_linear_persistent<true, 128>(mx, mx, whi, wlo, mout, maux, a, s);
}

// ---- weight / bias gradient -----------------------------------------------------------------------
static inline int wgrad_bn(int k_in) { return k_in <= 32 ? 32 : (k_in <= 64 ? 64 : 128); }

static int wgrad_splits(int64_t rows, int tiles) {
  // two CTAs fit on an SM: fill at most two whole waves (rounding the split count UP would leave a third,
  // nearly empty wave: 19 splits x 32 tiles = 608 CTAs on 592 slots cost 50% more than 18 x 32)
  int64_t want = (4 * kSMs) / tiles;
  int64_t max_splits = (rows + WG_ROWS - 1) / WG_ROWS;
  if (want > max_splits) want = max_splits;
  return (int)(want < 1 ? 1 : want);
}

extern "C" size_t t2h_linear_wgrad_workspace_bytes(int64_t rows, int n_out, int k_in) {
  if (rows <= 0 || n_out <= 0 || k_in <= 0) return 256;
  const int bn = wgrad_bn(k_in);
  const int tiles = ((n_out + BLOCK_M - 1) / BLOCK_M) * ((k_in + bn - 1) / bn);
  return (size_t)wgrad_splits(rows, tiles) * n_out * (k_in + 1) * sizeof(float) + 256;
}

template <int BLOCK_N, bool F16 = false>
static int launch_wgrad(const CUtensorMap& mg, const CUtensorMap& mx, WgradArgs a, int splits, cudaStream_t stream) {
  auto kern = wgrad_x3_kernel<BLOCK_N, F16>;
  // per device and idempotent; set on every launch so that the entry point keeps no state
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<BLOCK_N, F16>::TOTAL) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  dim3 grid((unsigned)((a.n_out + BLOCK_M - 1) / BLOCK_M), (unsigned)((a.k_in + BLOCK_N - 1) / BLOCK_N), (unsigned)splits);
  kern<<<grid, WG_THREADS, WgSmem<BLOCK_N, F16>::TOTAL, stream>>>(mg, mx, a);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_linear_wgrad(const float* grad_out, int64_t ld_g, const float* x, int64_t ld_x, int64_t rows,
                                int n_out, int k_in, int relu_in, void* workspace, size_t workspace_bytes,
                                float* grad_w, int64_t ld_w, float* grad_b, t2h_stream_t stream) {
  if (!grad_out || !x || !grad_w || !workspace || rows < 0 || n_out <= 0 || k_in <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if ((n_out % 4) || (k_in % 4) || (ld_g % 4) || (ld_x % 4) || (ld_w % 4)) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)grad_out | (uintptr_t)x | (uintptr_t)grad_w | (uintptr_t)workspace) & 15) return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_linear_wgrad_workspace_bytes(rows, n_out, k_in)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bn = wgrad_bn(k_in);
  const int tiles = ((n_out + BLOCK_M - 1) / BLOCK_M) * ((k_in + bn - 1) / bn);
  const int splits = rows > 0 ? wgrad_splits(rows, tiles) : 1;
  WgradArgs a;
  a.rows = rows; a.n_out = n_out; a.k_in = k_in; a.relu_in = relu_in;
  int64_t per = (rows + splits - 1) / splits;
  a.rows_per_split = ((per + WG_ROWS - 1) / WG_ROWS) * WG_ROWS;
  if (a.rows_per_split < WG_ROWS) a.rows_per_split = WG_ROWS;
  a.partial = (float*)workspace;
  a.partial_bias = grad_b ? a.partial + (size_t)splits * n_out * k_in : nullptr;
  a.conv = 0; a.tiles_x = a.units_per_img = a.cin = 0;
  a.g_absmax = a.x_absmax = nullptr;
  a.a_cols = n_out >= BLOCK_M ? BLOCK_M : ((n_out + 31) / 32) * 32;
  int st = T2H_OK;
  if (rows > 0) {
    CUtensorMap mg, mx;
    if (!make_map(&mg, grad_out, n_out, rows, ld_g, a.a_cols, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    if (!make_map(&mx, x, k_in, rows, ld_x, 32, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return T2H_ERR_CUDA;
    if (bn == 32) st = launch_wgrad<32>(mg, mx, a, splits, s);
    else if (bn == 64) st = launch_wgrad<64>(mg, mx, a, splits, s);
    else st = launch_wgrad<128>(mg, mx, a, splits, s);
    if (st) return st;
  }
  launch_wgrad_reduce(a.partial, a.partial_bias, rows > 0 ? splits : 0, n_out, k_in, grad_w, ld_w, grad_b, s);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}


extern "C" int t2h_linear_wgrad_f16(const float* grad_out, int64_t ld_g, const uint32_t* g_absmax, const float* x,
                                    int64_t ld_x, const uint32_t* x_absmax, int64_t rows, int n_out, int k_in, int relu_in,
                                    void* workspace, size_t workspace_bytes, float* grad_w, int64_t ld_w, float* grad_b,
                                    t2h_stream_t stream) {
  if (!grad_out || !x || !g_absmax || !x_absmax || !grad_w || !workspace || rows < 0 || n_out <= 0 || k_in <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if ((n_out % 4) || (k_in % 4) || (ld_g % 4) || (ld_x % 4) || (ld_w % 4) || k_in <= 64) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)grad_out | (uintptr_t)x | (uintptr_t)grad_w | (uintptr_t)workspace) & 15) return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_linear_wgrad_workspace_bytes(rows, n_out, k_in)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles = ((n_out + BLOCK_M - 1) / BLOCK_M) * ((k_in + 127) / 128);
  const int splits = rows > 0 ? wgrad_splits(rows, tiles) : 1;
  WgradArgs a;
  a.rows = rows; a.n_out = n_out; a.k_in = k_in; a.relu_in = relu_in;
  int64_t per = (rows + splits - 1) / splits;
  a.rows_per_split = ((per + WG_ROWS - 1) / WG_ROWS) * WG_ROWS;
  if (a.rows_per_split < WG_ROWS) a.rows_per_split = WG_ROWS;
  a.partial = (float*)workspace;
  a.partial_bias = grad_b ? a.partial + (size_t)splits * n_out * k_in : nullptr;
  a.conv = 0; a.tiles_x = a.units_per_img = a.cin = 0;
  a.g_absmax = g_absmax; a.x_absmax = x_absmax;
  a.a_cols = n_out >= BLOCK_M ? BLOCK_M : ((n_out + 31) / 32) * 32;
  if (rows > 0 && n_out % WGP_TILE == 0 && k_in % WGP_TILE == 0 && ablation_switch("T2H_WGRAD_PAIR", 1)) {
    // wide layers: CTA pairs on 256 x 256 tiles, one wave of pairs (never more splits than the workspace holds)
    const int tiles_k = k_in / WGP_TILE, n_tiles = (n_out / WGP_TILE) * tiles_k;
    int64_t want = (kSMs / 2) / n_tiles;
    if (want < 1) want = 1;
    if (want > splits) want = splits;
    per = (rows + want - 1) / want;
    a.rows_per_split = ((per + WG_ROWS - 1) / WG_ROWS) * WG_ROWS;
    const int psplits = (int)((rows + a.rows_per_split - 1) / a.rows_per_split);  // every split holds rows
    a.partial_bias = grad_b ? a.partial + (size_t)psplits * n_out * k_in : nullptr;
    CUtensorMap mg, mx;
    if (!make_map(&mg, grad_out, n_out, rows, ld_g, 128, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    if (!make_map(&mx, x, k_in, rows, ld_x, 32, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    if (cudaFuncSetAttribute(wgrad_f16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGP_TOTAL) != cudaSuccess) {
      (void)cudaGetLastError();
      return T2H_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * n_tiles * psplits));
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = WGP_TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, wgrad_f16_pair_kernel, mg, mx, a, tiles_k, n_tiles) != cudaSuccess) {
      (void)cudaGetLastError();
      return T2H_ERR_CUDA;
    }
    launch_wgrad_reduce(a.partial, a.partial_bias, psplits, n_out, k_in, grad_w, ld_w, grad_b, s);
    T2H_CHECK_LAUNCH();
    return T2H_OK;
  }
  if (rows > 0) {
    CUtensorMap mg, mx;
    if (!make_map(&mg, grad_out, n_out, rows, ld_g, a.a_cols, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    if (!make_map(&mx, x, k_in, rows, ld_x, 32, WG_ROWS, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    const int st = launch_wgrad<128, true>(mg, mx, a, splits, s);
    if (st) return st;
  }
  launch_wgrad_reduce(a.partial, a.partial_bias, rows > 0 ? splits : 0, n_out, k_in, grad_w, ld_w, grad_b, s);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" size_t t2h_conv3x3_wgrad_workspace_bytes(int B, int H, int W, int cin, int cout) {
  return t2h_linear_wgrad_workspace_bytes((int64_t)B * H * W, cout, 9 * cin);
}

static int conv3x3_wgrad_impl(const float* grad_out, const float* x, int B, int H, int W, int cin, int cout,
                              int relu_in, void* workspace, size_t workspace_bytes, float* grad_w, float* grad_b,
                              const uint32_t* g_absmax, const uint32_t* x_absmax, t2h_stream_t stream) {
  const bool f16 = g_absmax != nullptr;
  if (!grad_out || !x || !grad_w || !workspace || B < 0 || H <= 0 || W <= 0 || cin <= 0 || cout <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if ((cin % 32) || (cout % 4) || (W % 16) || (H % 2)) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)grad_out | (uintptr_t)x | (uintptr_t)grad_w | (uintptr_t)workspace) & 15) return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_conv3x3_wgrad_workspace_bytes(B, H, W, cin, cout)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t rows = (int64_t)B * H * W;
  const int k_in = 9 * cin;
  const int bn = f16 ? 128 : wgrad_bn(k_in);
  const int tiles = ((cout + BLOCK_M - 1) / BLOCK_M) * ((k_in + bn - 1) / bn);
  const int splits = rows > 0 ? wgrad_splits(rows, tiles) : 1;
  WgradArgs a;
  a.rows = rows; a.n_out = cout; a.k_in = k_in; a.relu_in = relu_in;
  int64_t per = (rows + splits - 1) / splits;
  a.rows_per_split = ((per + WG_ROWS - 1) / WG_ROWS) * WG_ROWS;
  if (a.rows_per_split < WG_ROWS) a.rows_per_split = WG_ROWS;
  a.partial = (float*)workspace;
  a.partial_bias = grad_b ? a.partial + (size_t)splits * cout * k_in : nullptr;
  a.conv = 1; a.tiles_x = W / 16; a.units_per_img = (H / 2) * a.tiles_x; a.cin = cin;
  a.g_absmax = g_absmax; a.x_absmax = x_absmax;
  a.a_cols = cout >= BLOCK_M ? BLOCK_M : ((cout + 31) / 32) * 32;
  if (rows > 0) {
    CUtensorMap mg, mx;
    if (!make_map_4d(&mg, grad_out, cout, W, H, B, a.a_cols, 16, 2, CU_TENSOR_MAP_SWIZZLE_NONE)) return T2H_ERR_CUDA;
    if (!make_map_4d(&mx, x, cin, W, H, B, 32, 16, 2, f16 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return T2H_ERR_CUDA;
    int st;
    if (f16) st = launch_wgrad<128, true>(mg, mx, a, splits, s);
    else if (bn == 32) st = launch_wgrad<32>(mg, mx, a, splits, s);
    else if (bn == 64) st = launch_wgrad<64>(mg, mx, a, splits, s);
    else st = launch_wgrad<128>(mg, mx, a, splits, s);
    if (st) return st;
  }
  launch_wgrad_reduce(a.partial, a.partial_bias, rows > 0 ? splits : 0, cout, k_in, grad_w, k_in, grad_b, s);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_conv3x3_wgrad(const float* grad_out, const float* x, int B, int H, int W, int cin, int cout,
                                 int relu_in, void* workspace, size_t workspace_bytes, float* grad_w, float* grad_b,
                                 t2h_stream_t stream) {
  return conv3x3_wgrad_impl(grad_out, x, B, H, W, cin, cout, relu_in, workspace, workspace_bytes, grad_w, grad_b, nullptr,
                            nullptr, stream);
}

extern "C" int t2h_conv3x3_wgrad_f16(const float* grad_out, const uint32_t* g_absmax, const float* x, const uint32_t* x_absmax,
                                     int B, int H, int W, int cin, int cout, int relu_in, void* workspace,
                                     size_t workspace_bytes, float* grad_w, float* grad_b, t2h_stream_t stream) {
  if (!g_absmax || !x_absmax) return T2H_ERR_INVALID_ARGUMENT;
  return conv3x3_wgrad_impl(grad_out, x, B, H, W, cin, cout, relu_in, workspace, workspace_bytes, grad_w, grad_b, g_absmax,
                            x_absmax, stream);
}

extern "C" size_t t2h_colsum_workspace_bytes(int64_t rows, int n) {
  (void)rows;
  return (size_t)4 * kSMs * (n > 0 ? n : 1) * sizeof(float) + 256;
}

extern "C" int t2h_colsum(const float* g, int64_t ld_g, int64_t rows, int n, void* workspace, size_t workspace_bytes,
                          float* out, t2h_stream_t stream) {
  if (!g || !out || !workspace || rows < 0 || n <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_colsum_workspace_bytes(rows, n)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  int blocks = 4 * kSMs;
  int64_t rpb = (rows + blocks - 1) / blocks;
  if (rpb < 8) rpb = 8;
  blocks = (int)((rows + rpb - 1) / rpb);
  float* partial = (float*)workspace;
  if (blocks > 0) {
    if (n <= 32) colsum_partial_kernel<32><<<dim3(blocks, 1), 256, 0, s>>>(g, rows, n, ld_g, rpb, partial);
    else if (n <= 64) colsum_partial_kernel<64><<<dim3(blocks, 1), 256, 0, s>>>(g, rows, n, ld_g, rpb, partial);
    else if (n <= 128) colsum_partial_kernel<128><<<dim3(blocks, 1), 256, 0, s>>>(g, rows, n, ld_g, rpb, partial);
    else colsum_partial_kernel<256><<<dim3(blocks, 1), 256, 0, s>>>(g, rows, n, ld_g, rpb, partial);
    T2H_CHECK_LAUNCH();
  }
  colsum_final_kernel<<<(n + 255) / 256, 256, 0, s>>>(partial, blocks, n, out);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}
