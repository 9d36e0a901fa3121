// Bilinear plane -> point sampling and plane up-sampling, forward and atomic-free backward
// (G1 / G2 / G3 of SURVEY §2.2).
//
// Replace F.grid_sample(..., mode='bilinear', padding_mode='border', align_corners=True)
// (alto.py:90-95, alto.py:199-205) and F.interpolate(..., mode='bilinear', align_corners=True)
// (pixel.py:105-111) together with their ATen backward kernels, which scatter the four corner
// contributions with atomicAdd.  Here points are visited in cell-sorted (Morton) order so the
// four taps of neighbouring points hit the same L1 lines, and the backward is a GATHER: every
// plane cell sums the contributions of the points in its 3x3 cell neighbourhood in a fixed
// order (grid_sample taps of a point in cell cx are always within {cx-1, cx, cx+1}, because the
// corner-aligned coordinate p*(r-1) lies in (cx-1, cx+1) when p*r is in [cx, cx+1)).
#include "t2h_common.cuh"
#include <cstdlib>

namespace t2h {

constexpr int kSampleWarps = 8;
constexpr int kPointsPerWarp = 32;  // sorted points handled by one warp in the forward

struct Taps {
  int x0, y0;
  float wx0, wx1, wy0, wy1;  // weight of column x0 / x0+1 and row y0 / y0+1
};

__device__ __forceinline__ Taps make_taps(float px, float py, int reso) {
  Taps t;
  const float ix = unnormalize_border(px, reso);
  const float iy = unnormalize_border(py, reso);
  const float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  t.wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), ix);  // ix_se - ix
  t.wx1 = __fsub_rn(ix, fx);                   // ix - ix_nw
  t.wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), iy);
  t.wy1 = __fsub_rn(iy, fy);
  return t;
}

// Forward: a warp owns 32 consecutive sorted points.  Lane j computes the taps of point j ONCE (the
// coordinate arithmetic is ~40 instructions and these kernels are issue-bound, not bandwidth-bound);
// the RPI sub-groups then walk the points and fetch (offset, weights, row) by shuffle.
template <class RS>
__global__ void __launch_bounds__(kSampleWarps * kWarp)
sample_fwd_kernel(const float* __restrict__ plane, int reso, const float* __restrict__ xyz, int64_t stride,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_ids, int64_t n, int64_t n_per_batch,
                  float* __restrict__ out) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t warp = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
  const int64_t first = warp * kPointsPerWarp;
  if (first >= n) return;
  const int npts = (int)min((int64_t)kPointsPerWarp, n - first);

  // ---- per-lane tap setup for point first + lane -------------------------------------------------
  int64_t my_row = 0, my_off = 0;
  int my_dx = 0, my_dy = 0;
  float my_nw = 0.f, my_ne = 0.f, my_sw = 0.f, my_se = 0.f;
  if (lane < npts) {
    const int64_t i = first + lane;
    my_row = perm ? (int64_t)perm[i] : i;
    const int64_t b = tile_ids ? (int64_t)tile_ids[i] : my_row / n_per_batch;  // ragged batches carry tile ids
    const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + i * stride));
    const Taps t = make_taps(pxy.x, pxy.y, reso);
    // taps beyond the last row/column carry zero weight (ix == r-1 exactly); clamp the address
    my_dx = (t.x0 + 1 < reso) ? 1 : 0;
    my_dy = (t.y0 + 1 < reso) ? 1 : 0;
    my_nw = __fmul_rn(t.wx0, t.wy0);
    my_ne = my_dx ? __fmul_rn(t.wx1, t.wy0) : 0.f;
    my_sw = my_dy ? __fmul_rn(t.wx0, t.wy1) : 0.f;
    my_se = (my_dx && my_dy) ? __fmul_rn(t.wx1, t.wy1) : 0.f;
    my_off = (b * reso + t.y0) * (int64_t)reso + t.x0;  // pixel index of the north-west tap
  }
  const int my_flags = my_dx | (my_dy << 1);
  for (int t0 = 0; t0 < npts; t0 += RPI) {
    const int j = t0 + sub;
    const bool act = j < npts;
    const int src = act ? j : 0;
    const int64_t row = __shfl_sync(0xffffffffu, my_row, src);
    const int64_t off = __shfl_sync(0xffffffffu, my_off, src);
    const int flags = __shfl_sync(0xffffffffu, my_flags, src);
    const float w_nw = __shfl_sync(0xffffffffu, my_nw, src), w_ne = __shfl_sync(0xffffffffu, my_ne, src);
    const float w_sw = __shfl_sync(0xffffffffu, my_sw, src), w_se = __shfl_sync(0xffffffffu, my_se, src);
    if (!act) continue;
    const float* p_nw = plane + off * C + l * 4;
    const float* p_ne = p_nw + (flags & 1) * C;
    const float* p_sw = p_nw + (int64_t)(flags >> 1) * reso * C;
    const float* p_se = p_sw + (flags & 1) * C;
    float* dst = out + row * C + l * 4;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int o = c * LPR * 4;
      const float4 a = ld4(p_nw + o), bq = ld4(p_ne + o), cq = ld4(p_sw + o), d = ld4(p_se + o);
      float4 r;
      r.x = a.x * w_nw + bq.x * w_ne + cq.x * w_sw + d.x * w_se;
      r.y = a.y * w_nw + bq.y * w_ne + cq.y * w_sw + d.y * w_se;
      r.z = a.z * w_nw + bq.z * w_ne + cq.z * w_sw + d.z * w_se;
      r.w = a.w * w_nw + bq.w * w_ne + cq.w * w_sw + d.w * w_se;
      st4(dst + o, r);
    }
  }
}

// WPS warps per plane cell (cells enumerated in key order of this level, so a CTA covers a compact
// block of cells when keys are Morton codes).  Every warp scans a slice of each neighbour cell's
// points; slices are combined through shared memory in slice order (fixed summation order).
template <class RS, int WPS>
__global__ void __launch_bounds__(kSampleWarps * kWarp)
sample_bwd_kernel(const float* __restrict__ grad_rows, int reso, const float* __restrict__ xyz, int64_t stride,
                  const int32_t* __restrict__ perm, const int32_t* __restrict__ cell_start, int64_t n_seg, int shift,
                  int morton, int log2_cells, float* __restrict__ grad_plane) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  constexpr int SEGS = kSampleWarps / WPS;
  __shared__ float4 part_sum[WPS > 1 ? kSampleWarps * (C / 4) : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t seg = (int64_t)blockIdx.x * SEGS + warp / WPS;
  const int part = warp % WPS;
  const bool valid = seg < n_seg;
  const int64_t cells = (int64_t)reso * reso;
  // Morton levels have power-of-two cell counts: split tile / cell with shifts (no 64-bit division)
  const int64_t b = !valid ? 0 : (morton ? (seg >> log2_cells) : seg / cells);
  int cx = 0, cy = 0;
  if (valid) cell_decode((uint32_t)(seg - b * cells), reso, morton, cx, cy);
  // Morton bits of the three candidate columns / rows, computed once instead of per neighbour cell
  uint32_t mx[3], my[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mx[d] = part1by1((uint32_t)(cx + d - 1));
    my[d] = part1by1((uint32_t)(cy + d - 1)) << 1;
  }

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  if (valid) {
    for (int dy = -1; dy <= 1; ++dy) {
      const int ny = cy + dy;
      if (ny < 0 || ny >= reso) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int nx = cx + dx;
        if (nx < 0 || nx >= reso) continue;
        const int64_t key = b * cells + (morton ? (mx[dx + 1] | my[dy + 1]) : (uint32_t)(nx + reso * ny));
        int beg = cell_start[key << shift], end = cell_start[(key + 1) << shift];
        if (WPS > 1) {
          const int slice = (end - beg + WPS - 1) / WPS;
          beg = min(beg + part * slice, end);
          end = min(beg + slice, end);
        }
        // 32 candidates at a time: lane j evaluates the tap weights of candidate j once, the ballot keeps
        // only the contributing ones, and the RPI sub-groups share them out in rank order (fixed order)
        for (int base_i = beg; base_i < end; base_i += kWarp) {
          const int i = base_i + lane;
          float w = 0.f;
          int row = 0;
          if (i < end) {
            const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + (int64_t)i * stride));
            const Taps t = make_taps(pxy.x, pxy.y, reso);
            const float wx = (t.x0 == cx) ? t.wx0 : ((t.x0 + 1 == cx) ? t.wx1 : 0.f);
            const float wy = (t.y0 == cy) ? t.wy0 : ((t.y0 + 1 == cy) ? t.wy1 : 0.f);
            w = __fmul_rn(wx, wy);
            row = perm ? perm[i] : i;
          }
          const unsigned hit = __ballot_sync(0xffffffffu, w != 0.f);
          const int cnt = __popc(hit);
          for (int t0 = 0; t0 < cnt; t0 += RPI) {
            const int k = t0 + sub;
            const bool act = k < cnt;
            const int src = act ? (int)__fns(hit, 0, k + 1) : 0;
            const float wk = __shfl_sync(0xffffffffu, w, src);
            const int rk = __shfl_sync(0xffffffffu, row, src);
            if (act) {
              const float* srcp = grad_rows + (int64_t)rk * C + l * 4;
#pragma unroll
              for (int c = 0; c < CH; ++c) {
                const float4 gq = ld4(srcp + c * LPR * 4);
                acc[c].x += wk * gq.x; acc[c].y += wk * gq.y; acc[c].z += wk * gq.z; acc[c].w += wk * gq.w;
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int off = LPR; off < kWarp; off <<= 1)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 o = shfl_xor4(acc[c], off);
      acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
    }
  if (WPS > 1) {
    if (sub == 0 && part > 0) {
#pragma unroll
      for (int c = 0; c < CH; ++c) part_sum[warp * (C / 4) + c * LPR + l] = acc[c];
    }
    __syncthreads();
    if (part == 0 && sub == 0) {
      for (int q = 1; q < WPS; ++q)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float4 o = part_sum[(warp + q) * (C / 4) + c * LPR + l];
          acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
        }
    }
  }
  if (valid && part == 0 && sub == 0) {
    float* dst = grad_plane + ((b * reso + cy) * (int64_t)reso + cx) * C + l * 4;
#pragma unroll
    for (int c = 0; c < CH; ++c) st4(dst + c * LPR * 4, acc[c]);
  }
}

// ---- G2, tiled: shared-memory-staged, atomic-free, deterministic ------------------------------------
// A CTA owns a T x T block of plane cells.  With Morton keys its points are ONE contiguous range of the
// sorted order, so their gradient rows stream in coalesced and are read exactly once.  Every point
// scatters its four weighted contributions into a (T+2) x (T+2) accumulator tile in shared memory
// (taps of a point of cell cx lie in columns cx-1..cx+1, hence the one-cell halo).  Determinism without
// atomics comes from ownership: the 8 warps form CG channel groups x PG point groups; a warp only ever
// touches its own 32*V-channel slice of its point group's private tile, in point order, so each
// accumulator has a single writer and a fixed summation order.  Private tiles are then summed in group
// order into a per-block scratch tile; sample_bwd_merge_kernel adds the (<= 4) overlapping block tiles
// of every plane cell in a fixed order.
constexpr int kTileWarps = 8;

template <int C, int T>
struct TileCfg {
  static constexpr int CG = (C / 32 < kTileWarps) ? C / 32 : kTileWarps;  // channel groups
  static constexpr int V = C / (32 * CG);                                // floats per lane
  static constexpr int PG = kTileWarps / CG;                             // point groups (private tiles)
  static constexpr int TW = T + 2;
  static constexpr int TILE_FLOATS = TW * TW * C;
  static constexpr int SMEM = PG * TILE_FLOATS * 4;
};

template <int C, int T>
__global__ void __launch_bounds__(kTileWarps * kWarp)
sample_bwd_tiled_kernel(const float* __restrict__ grad_rows, int reso, const float* __restrict__ xyz, int64_t stride,
                        const int32_t* __restrict__ perm, const int32_t* __restrict__ cell_start, int shift,
                        int log2_cells, float* __restrict__ scratch) {
  using Cfg = TileCfg<C, T>;
  constexpr int CG = Cfg::CG, V = Cfg::V, PG = Cfg::PG, TW = Cfg::TW;
  extern __shared__ float tile_smem[];
  constexpr int LOG2_T2 = (T == 8) ? 6 : ((T == 4) ? 4 : ((T == 2) ? 2 : 0));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pg = warp / CG, cg = warp % CG;
  // block index -> tile (image) and Morton block code -> origin
  const int64_t blk = blockIdx.x;
  const int blocks_log2 = log2_cells - LOG2_T2;
  const int64_t img = blk >> blocks_log2;
  const uint32_t bcode = (uint32_t)(blk - (img << blocks_log2));
  const int bx0 = (int)compact1by1(bcode) * T, by0 = (int)compact1by1(bcode >> 1) * T;
  const int64_t key0 = (img << log2_cells) + ((int64_t)bcode << LOG2_T2);
  const int p0 = cell_start[key0 << shift], p1 = cell_start[(key0 + (1 << LOG2_T2)) << shift];

  for (int f = threadIdx.x; f < PG * Cfg::TILE_FLOATS / 4; f += kTileWarps * kWarp)
    reinterpret_cast<float4*>(tile_smem)[f] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  // this point group's contiguous slice of the block's points
  const int len = p1 - p0, chunk = (len + PG - 1) / PG;
  const int q0 = min(p0 + pg * chunk, p1), q1 = min(q0 + chunk, p1);
  float* priv = tile_smem + pg * Cfg::TILE_FLOATS + cg * 32 * V + lane * V;
  const float* gbase = grad_rows + cg * 32 * V + lane * V;
  for (int base_i = q0; base_i < q1; base_i += kWarp) {
    const int nb = min(kWarp, q1 - base_i);
    int my_cell = 0, my_row = 0;
    float my_nw = 0.f, my_ne = 0.f, my_sw = 0.f, my_se = 0.f;
    if (lane < nb) {
      const int i = base_i + lane;
      const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + (int64_t)i * stride));
      const Taps t = make_taps(pxy.x, pxy.y, reso);
      my_nw = __fmul_rn(t.wx0, t.wy0);
      my_ne = (t.x0 + 1 < reso) ? __fmul_rn(t.wx1, t.wy0) : 0.f;
      my_sw = (t.y0 + 1 < reso) ? __fmul_rn(t.wx0, t.wy1) : 0.f;
      my_se = (t.x0 + 1 < reso && t.y0 + 1 < reso) ? __fmul_rn(t.wx1, t.wy1) : 0.f;
      // local coordinates in the haloed tile; clamp defensively (a point is always within one cell)
      const int lx = min(max(t.x0 - bx0 + 1, 0), T), ly = min(max(t.y0 - by0 + 1, 0), T);
      my_cell = ly * TW + lx;
      my_row = perm ? perm[i] : i;
    }
    // U points per trip: issue all row loads first (memory-level parallelism), then the shared-memory updates
    constexpr int U = (V <= 2) ? 8 : 4;
    for (int j0 = 0; j0 < nb; j0 += U) {
      int cell[U];
      float wq[U][4], g[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = min(j0 + u, nb - 1);
        cell[u] = __shfl_sync(0xffffffffu, my_cell, j);
        const int row = __shfl_sync(0xffffffffu, my_row, j);
        wq[u][0] = __shfl_sync(0xffffffffu, my_nw, j); wq[u][1] = __shfl_sync(0xffffffffu, my_ne, j);
        wq[u][2] = __shfl_sync(0xffffffffu, my_sw, j); wq[u][3] = __shfl_sync(0xffffffffu, my_se, j);
        const float* gp = gbase + (int64_t)row * C;
        if constexpr (V == 1) {
          g[u][0] = __ldg(gp);
        } else if constexpr (V == 2) {
          const float2 t2 = __ldg(reinterpret_cast<const float2*>(gp));
          g[u][0] = t2.x; g[u][1] = t2.y;
        } else {
          const float4 t4 = ld4(gp);
          g[u][0] = t4.x; g[u][1] = t4.y; g[u][2] = t4.z; g[u][3] = t4.w;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u >= nb) break;
        float* a = priv + cell[u] * C;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          a[v] += wq[u][0] * g[u][v];
          a[C + v] += wq[u][1] * g[u][v];
          a[TW * C + v] += wq[u][2] * g[u][v];
          a[TW * C + C + v] += wq[u][3] * g[u][v];
        }
      }
    }
  }
  __syncthreads();
  // private tiles -> scratch, summed in point-group order
  float* dst = scratch + blk * (int64_t)Cfg::TILE_FLOATS;
  for (int f = threadIdx.x; f < Cfg::TILE_FLOATS / 4; f += kTileWarps * kWarp) {
    float4 acc = reinterpret_cast<const float4*>(tile_smem)[f];
#pragma unroll
    for (int q = 1; q < PG; ++q) {
      const float4 o = reinterpret_cast<const float4*>(tile_smem + q * Cfg::TILE_FLOATS)[f];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    st4(dst + f * 4, acc);
  }
}

// ---- G2, cell-parallel: register accumulation of the 3x3 target partials ------------------------------
// All points of one own cell (cx, cy) contribute to the same 3x3 target cells (cx-1..cx+1, cy-1..cy+1),
// so a warp that walks a cell's (contiguous) rows keeps nine partial sums in REGISTERS -- no per-point
// shared-memory traffic, rows are read once, coalesced.  A CTA owns a T x T Morton block of cells; per
// round its 8 warps are split into (cells x row slices x 128-channel groups).  Slice partials are combined
// through shared memory in slice order, every own cell's nine partials are staged in shared memory, and
// the block's haloed (T+2) x (T+2) output tile is assembled from them in a fixed order and written to
// scratch; sample_bwd_merge_kernel adds the overlapping block tiles.  No atomics, fixed summation order.
template <int C, int T>
struct CellCfg {
  static constexpr int CW = C < 128 ? C : 128;   // channels per warp
  static constexpr int CGN = C / CW;             // channel groups
  static constexpr int RF = kTileWarps / CGN;    // warps per channel group = cells per round x slices
  static constexpr int LPR = CW / 4, RPI = 32 / LPR;
  static constexpr int NC = T * T, TW = T + 2;
  static constexpr int PARTS_FLOATS = kTileWarps * 9 * CW;
  static constexpr int STAGE_FLOATS = NC * 9 * C;
  static constexpr int SMEM = (PARTS_FLOATS + STAGE_FLOATS) * 4;
  static constexpr int LOG2_T2 = (T == 8) ? 6 : ((T == 4) ? 4 : ((T == 2) ? 2 : 0));
};

template <int C, int T>
__global__ void __launch_bounds__(kTileWarps * kWarp)
sample_bwd_cell_kernel(const float* __restrict__ grad_rows, int reso, const float* __restrict__ xyz, int64_t stride,
                       const int32_t* __restrict__ perm, const int32_t* __restrict__ cell_start, int shift,
                       int log2_cells, int slices, float* __restrict__ scratch) {
  // gridDim.y = ZS: on coarse levels (hundreds of rows per cell) the rows of every cell are additionally
  // split over ZS CTAs, each writing its own scratch tile; the merge kernel adds them in z order
  using Cfg = CellCfg<C, T>;
  constexpr int CW = Cfg::CW, CGN = Cfg::CGN, RF = Cfg::RF, LPR = Cfg::LPR, RPI = Cfg::RPI, NC = Cfg::NC, TW = Cfg::TW;
  extern __shared__ float cell_smem[];
  float* parts = cell_smem;                       // [warp][9][CW]
  float* stage = cell_smem + Cfg::PARTS_FLOATS;   // [own cell][9][C]
  __shared__ int cell_heavy[NC];
  constexpr int kHeavyRows = 96;                  // rows per slice beyond which a cell counts as heavy
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int cg = warp % CGN, rest = warp / CGN;
  const int sl = rest % slices, cr = rest / slices;
  const int cells_per_round = RF / slices;

  const int64_t blk = blockIdx.x;
  const int blocks_log2 = log2_cells - Cfg::LOG2_T2;
  const int64_t img = blk >> blocks_log2;
  const uint32_t bcode = (uint32_t)(blk - (img << blocks_log2));
  const int bx0 = (int)compact1by1(bcode) * T, by0 = (int)compact1by1(bcode >> 1) * T;
  const int64_t key0 = (img << log2_cells) + ((int64_t)bcode << Cfg::LOG2_T2);
  const float* gbase = grad_rows + cg * CW + l * 4;

  // rows of own cell q, slice my_sl of n_sl  ->  nine partial sums in registers  ->  parts[warp]
  auto reduce_cell = [&](int q, int my_sl, int n_sl) {
    float4 acc[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int cx = bx0 + (int)compact1by1((uint32_t)q), cy = by0 + (int)compact1by1((uint32_t)q >> 1);
    const int beg = cell_start[(key0 + q) << shift], end = cell_start[(key0 + q + 1) << shift];
    const int zs = gridDim.y, tot_sl = n_sl * zs, g_sl = blockIdx.y * n_sl + my_sl;
    const int len = end - beg, chunk = (len + tot_sl - 1) / tot_sl;
    const int r0 = min(beg + g_sl * chunk, end), r1 = min(r0 + chunk, end);
    for (int base_i = r0; base_i < r1; base_i += kWarp) {
      const int nb = min(kWarp, r1 - base_i);
      // lane j: the 3 + 3 separable target weights of row base_i + j (the tap arithmetic runs once per row)
      float wx_m = 0.f, wx_0 = 0.f, wx_p = 0.f, wy_m = 0.f, wy_0 = 0.f, wy_p = 0.f;
      int my_row = 0;
      if (lane < nb) {
        const int i = base_i + lane;
        const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + (int64_t)i * stride));
        const Taps t = make_taps(pxy.x, pxy.y, reso);
        const float wx1 = (t.x0 + 1 < reso) ? t.wx1 : 0.f, wy1 = (t.y0 + 1 < reso) ? t.wy1 : 0.f;
        if (t.x0 < cx) { wx_m = t.wx0; wx_0 = wx1; } else { wx_0 = t.wx0; wx_p = wx1; }
        if (t.y0 < cy) { wy_m = t.wy0; wy_0 = wy1; } else { wy_0 = t.wy0; wy_p = wy1; }
        my_row = perm ? perm[i] : i;
      }
      // U rows per sub-group per trip: all U row loads are issued before the FMAs (memory-level parallelism)
      constexpr int U = (RPI > 1) ? 1 : 4;  // narrow rows already cover RPI rows per trip; fine cells hold ~4 rows
      for (int t0 = 0; t0 < nb; t0 += RPI * U) {
        float4 g[U];
        float wx[U][3], wy[U][3];
        bool act[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = t0 + u * RPI + sub;
          act[u] = j < nb;
          const int src = act[u] ? j : 0;
          const int row = __shfl_sync(0xffffffffu, my_row, src);
          wx[u][0] = __shfl_sync(0xffffffffu, wx_m, src); wx[u][1] = __shfl_sync(0xffffffffu, wx_0, src); wx[u][2] = __shfl_sync(0xffffffffu, wx_p, src);
          wy[u][0] = __shfl_sync(0xffffffffu, wy_m, src); wy[u][1] = __shfl_sync(0xffffffffu, wy_0, src); wy[u][2] = __shfl_sync(0xffffffffu, wy_p, src);
          g[u] = act[u] ? ld4(gbase + (int64_t)row * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (!act[u]) continue;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float w = __fmul_rn(wx[u][dx], wy[u][dy]);
              float4& a = acc[dy * 3 + dx];
              a.x += w * g[u].x; a.y += w * g[u].y; a.z += w * g[u].z; a.w += w * g[u].w;
            }
        }
      }
    }
    // fold the RPI sub-rows of the warp (fixed xor tree)
#pragma unroll
    for (int off = LPR; off < kWarp; off <<= 1)
#pragma unroll
      for (int d = 0; d < 9; ++d) {
        const float4 o = shfl_xor4(acc[d], off);
        acc[d].x += o.x; acc[d].y += o.y; acc[d].z += o.z; acc[d].w += o.w;
      }
    if (sub == 0) {
#pragma unroll
      for (int d = 0; d < 9; ++d) *reinterpret_cast<float4*>(parts + (warp * 9 + d) * CW + l * 4) = acc[d];
    }
  };
  // parts of n_cells own cells (n_sl slices each) -> stage, in slice order; cells flagged `skip_heavy` are left out
  auto combine = [&](int q_first, int n_cells, int n_sl, bool skip_heavy) {
    for (int idx = threadIdx.x; idx < n_cells * 9 * (C / 4); idx += kTileWarps * kWarp) {
      const int c4i = idx % (C / 4);
      const int d = (idx / (C / 4)) % 9;
      const int crr = idx / (9 * (C / 4));
      if (skip_heavy && cell_heavy[q_first + crr]) continue;
      const int cgi = (c4i * 4) / CW, within = (c4i * 4) % CW;
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s2 = 0; s2 < n_sl; ++s2) {
        const int w2 = (crr * n_sl + s2) * CGN + cgi;
        const float4 o = *reinterpret_cast<const float4*>(parts + (w2 * 9 + d) * CW + within);
        sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
      }
      *reinterpret_cast<float4*>(stage + ((q_first + crr) * 9 + d) * C + c4i * 4) = sum;
    }
  };

  // heavy cells (a facade crossing the block) are deferred and then reduced by every warp of the CTA
  if (threadIdx.x < NC) {
    const int len = cell_start[(key0 + threadIdx.x + 1) << shift] - cell_start[(key0 + threadIdx.x) << shift];
    cell_heavy[threadIdx.x] = (slices < RF && len > kHeavyRows * slices * (int)gridDim.y) ? 1 : 0;
  }
  __syncthreads();
  for (int q0 = 0; q0 < NC; q0 += cells_per_round) {
    const int q = q0 + cr;
    if (q < NC && !cell_heavy[q]) reduce_cell(q, sl, slices);
    __syncthreads();
    combine(q0, min(cells_per_round, NC - q0), slices, true);
    __syncthreads();
  }
  for (int q = 0; q < NC; ++q) {
    if (!cell_heavy[q]) continue;  // uniform across the CTA
    reduce_cell(q, rest, RF);
    __syncthreads();
    combine(q, 1, RF, false);
    __syncthreads();
  }
  // haloed output tile of the block: target (ly, lx) collects partial d = (dy, dx) of own cell (ly-1-dy, lx-1-dx)
  float* dst = scratch + ((int64_t)blockIdx.y * gridDim.x + blk) * (int64_t)(TW * TW * C);
  for (int idx = threadIdx.x; idx < TW * TW * (C / 4); idx += kTileWarps * kWarp) {
    const int c4i = idx % (C / 4), cellt = idx / (C / 4);
    const int ly = cellt / TW, lx = cellt % TW;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int oy = ly - dy, ox = lx - dx;  // = (ly - 1) - (dy - 1)
        if (ox >= 0 && ox < T && oy >= 0 && oy < T) {
          const int qq = (int)(part1by1((uint32_t)ox) | (part1by1((uint32_t)oy) << 1));
          const float4 o = *reinterpret_cast<const float4*>(stage + (qq * 9 + dy * 3 + dx) * C + c4i * 4);
          sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
        }
      }
    st4(dst + (int64_t)cellt * C + c4i * 4, sum);
  }
}

// grad_plane[b, y, x, :] = sum of the block tiles that cover (x, y): own block plus the left/upper or
// right/lower neighbours when the cell lies on a block edge; fixed order (by, then bx)
template <int T>
__global__ void __launch_bounds__(256)
sample_bwd_merge_kernel(const float* __restrict__ scratch, int reso, int C, int log2_cells, int64_t n_cells, int zs,
                        float* __restrict__ grad_plane) {
  constexpr int TW = T + 2;
  constexpr int LOG2_T2 = (T == 8) ? 6 : ((T == 4) ? 4 : ((T == 2) ? 2 : 0));
  const int c4 = C / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * c4) return;
  const int64_t cellid = gid / c4;
  const int ch = (int)(gid - cellid * c4) * 4;
  const int64_t cells = (int64_t)reso * reso;
  const int64_t img = cellid / cells;
  const int rem = (int)(cellid - img * cells);
  const int y = rem / reso, x = rem - y * reso;
  const int nblk = reso / T;
  const int blocks_log2 = log2_cells - LOG2_T2;
  const int bx_lo = max((x - 1 + T) / T - 1, 0) , bx_hi = min((x + 1) / T, nblk - 1);
  const int by_lo = max((y - 1 + T) / T - 1, 0) , by_hi = min((y + 1) / T, nblk - 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n_blocks = n_cells >> LOG2_T2;
  for (int z = 0; z < zs; ++z)
    for (int by = by_lo; by <= by_hi; ++by)
      for (int bx = bx_lo; bx <= bx_hi; ++bx) {
        const int lx = x - bx * T + 1, ly = y - by * T + 1;
        if (lx < 0 || lx >= TW || ly < 0 || ly >= TW) continue;
        const int64_t blk = z * n_blocks + (img << blocks_log2) + (part1by1((uint32_t)bx) | (part1by1((uint32_t)by) << 1));
        const float4 v = ld4(scratch + (blk * (TW * TW) + ly * TW + lx) * (int64_t)C + ch);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
  st4(grad_plane + cellid * C + ch, acc);
}

// ---- regular-grid bilinear resize, align_corners=True (ATen UpSampleBilinear2d semantics) -----
struct Axis {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Axis make_axis(int dst, float scale, int n_in) {
  Axis a;
  const float src = __fmul_rn(scale, (float)dst);
  a.i0 = min((int)src, n_in - 1);
  a.i1 = a.i0 + ((a.i0 < n_in - 1) ? 1 : 0);
  a.l1 = __fsub_rn(src, (float)a.i0);
  a.l0 = __fsub_rn(1.0f, a.l1);
  return a;
}

template <class RS>
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float* __restrict__ in, int B, int h, int w, int oh, int ow, float sh, float sw,
                    float* __restrict__ out) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t total = (int64_t)B * oh * ow;
  const int64_t pix = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPI + sub;
  if (pix >= total) return;
  const int ox = (int)(pix % ow);
  const int oy = (int)((pix / ow) % oh);
  const int64_t b = pix / ((int64_t)ow * oh);
  const Axis ax = make_axis(ox, sw, w), ay = make_axis(oy, sh, h);
  const float* base = in + b * h * (int64_t)w * C + l * 4;
  const float* p00 = base + ((int64_t)ay.i0 * w + ax.i0) * C;
  const float* p01 = base + ((int64_t)ay.i0 * w + ax.i1) * C;
  const float* p10 = base + ((int64_t)ay.i1 * w + ax.i0) * C;
  const float* p11 = base + ((int64_t)ay.i1 * w + ax.i1) * C;
  float* dst = out + pix * C + l * 4;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int o = c * LPR * 4;
    const float4 a = ld4(p00 + o), bq = ld4(p01 + o), cq = ld4(p10 + o), d = ld4(p11 + o);
    float4 r;
    r.x = ay.l0 * (ax.l0 * a.x + ax.l1 * bq.x) + ay.l1 * (ax.l0 * cq.x + ax.l1 * d.x);
    r.y = ay.l0 * (ax.l0 * a.y + ax.l1 * bq.y) + ay.l1 * (ax.l0 * cq.y + ax.l1 * d.y);
    r.z = ay.l0 * (ax.l0 * a.z + ax.l1 * bq.z) + ay.l1 * (ax.l0 * cq.z + ax.l1 * d.z);
    r.w = ay.l0 * (ax.l0 * a.w + ax.l1 * bq.w) + ay.l1 * (ax.l0 * cq.w + ax.l1 * d.w);
    st4_stream(dst + o, r);
  }
}

// gather backward: input pixel (y, x) collects every output pixel that has it as one of its taps
template <class RS>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ grad_out, int B, int h, int w, int oh, int ow, float sh, float sw,
                    float inv_sh, float inv_sw, float* __restrict__ grad_in) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t total = (int64_t)B * h * w;
  const int64_t pix = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPI + sub;
  if (pix >= total) return;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const int64_t b = pix / ((int64_t)w * h);
  // candidate output range: source coordinate in (x-1, x+1)  (conservative by one pixel each side)
  const int ox_lo = max((int)floorf((float)(x - 1) * inv_sw) - 1, 0);
  const int ox_hi = min((int)ceilf((float)(x + 1) * inv_sw) + 1, ow - 1);
  const int oy_lo = max((int)floorf((float)(y - 1) * inv_sh) - 1, 0);
  const int oy_hi = min((int)ceilf((float)(y + 1) * inv_sh) + 1, oh - 1);
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const Axis ay = make_axis(oy, sh, h);
    const float wy = ((ay.i0 == y) ? ay.l0 : 0.f) + ((ay.i1 == y) ? ay.l1 : 0.f);
    if (wy == 0.f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const Axis ax = make_axis(ox, sw, w);
      const float wx = ((ax.i0 == x) ? ax.l0 : 0.f) + ((ax.i1 == x) ? ax.l1 : 0.f);
      if (wx == 0.f) continue;
      const float wgt = wy * wx;
      const float* src = grad_out + ((b * oh + oy) * (int64_t)ow + ox) * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 gq = ld4(src + c * LPR * 4);
        acc[c].x += wgt * gq.x; acc[c].y += wgt * gq.y; acc[c].z += wgt * gq.z; acc[c].w += wgt * gq.w;
      }
    }
  }
  float* dst = grad_in + pix * C + l * 4;
#pragma unroll
  for (int c = 0; c < CH; ++c) st4(dst + c * LPR * 4, acc[c]);
}

}  // namespace t2h

using namespace t2h;

extern "C" int t2h_bilinear_sample_fwd(const float* plane, int reso, int C, const float* xyz_sorted,
                                       int64_t point_stride, const int32_t* perm, const int32_t* tile_ids,
                                       int64_t n_points, int64_t n_per_batch, float* out_rows, t2h_stream_t stream) {
  if (!plane || !xyz_sorted || !out_rows || reso <= 0 || point_stride < 2 || (point_stride & 1) || n_points < 0 ||
      n_per_batch <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if (n_points == 0) return T2H_OK;
  const int64_t warps = (n_points + kPointsPerWarp - 1) / kPointsPerWarp;
  const unsigned blocks = (unsigned)((warps + kSampleWarps - 1) / kSampleWarps);
  T2H_DISPATCH_ROWSHAPE(C, sample_fwd_kernel<RS><<<blocks, kSampleWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               plane, reso, xyz_sorted, point_stride, perm, tile_ids, n_points, n_per_batch, out_rows));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

// Block edge T for the tiled backward: the largest of {8, 4, 2, 1} that keeps ~<= 512 points per block
// (coarse levels hold hundreds of points per cell; small blocks keep thousands of CTAs in flight)
static inline int tiled_T(int reso, int C, int morton, int64_t n_points, int64_t n_seg) {
  if (!morton || C < 32 || (C & (C - 1)) || C > 1024 || n_seg <= 0) return 0;   // C in {32, 64, ..., 1024}
  const int64_t avg = n_points / n_seg > 0 ? n_points / n_seg : 1;
  int T = C <= 512 ? 8 : 4;
  while (T > 1 && ((int64_t)T * T * avg > 512 || T > reso)) T >>= 1;
  return T;
}

extern "C" size_t t2h_bilinear_sample_bwd_workspace_bytes(int reso, int C, int64_t n_seg, int morton) {
  if (!morton || n_seg <= 0) return 256;
  // haloed block tiles: T = 1 blocks (C >= 512) carry a 3 x 3 tile per cell and up to 8 row splits;
  // T >= 2 blocks at most 4 tile cells per cell
  const size_t per_cell = C >= 512 ? 9 * 8 : (C >= 128 ? 4 * 8 : 36 * 8 / 16);
  return (size_t)n_seg * per_cell * C * sizeof(float) + 256;
}

static inline int cell_zsplit(int64_t avg, int slices) {
  int zs = 1;
  while (zs < 8 && avg >= (int64_t)48 * slices * zs) zs *= 2;
  return zs;
}

template <int C, int T>
static int launch_cell(const float* grad_rows, int reso, const float* xyz, int64_t stride, const int32_t* perm,
                       const int32_t* cell_start, int64_t n_points, int64_t n_seg, int shift, int log2_cells,
                       float* scratch, float* grad_plane, cudaStream_t s) {
  using Cfg = CellCfg<C, T>;
  auto kern = sample_bwd_cell_kernel<C, T>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
      (void)cudaGetLastError();
      return T2H_ERR_CUDA;
    }
    configured = true;
  }
  // ~32 rows per warp on average: row slices inside the CTA first (bounded by the warps of a channel group),
  // then ZS CTAs per block
  const int64_t avg = n_points / n_seg;
  int slices = 1;
  while (slices < Cfg::RF && avg >= 48 * slices) slices *= 2;
  const int zs = cell_zsplit(avg, slices);
  const int64_t blocks = n_seg / (T * T);
  kern<<<dim3((unsigned)blocks, (unsigned)zs), kTileWarps * kWarp, Cfg::SMEM, s>>>(grad_rows, reso, xyz, stride, perm, cell_start, shift, log2_cells, slices, scratch);
  T2H_CHECK_LAUNCH();
  const int64_t threads = n_seg * (C / 4);
  sample_bwd_merge_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(scratch, reso, C, log2_cells, n_seg, zs, grad_plane);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

template <int C, int T>
static int launch_tiled(const float* grad_rows, int reso, const float* xyz, int64_t stride, const int32_t* perm,
                        const int32_t* cell_start, int64_t n_seg, int shift, int log2_cells, float* scratch,
                        float* grad_plane, cudaStream_t s) {
  using Cfg = TileCfg<C, T>;
  auto kern = sample_bwd_tiled_kernel<C, T>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
      (void)cudaGetLastError();
      return T2H_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t blocks = n_seg / (T * T);
  kern<<<(unsigned)blocks, kTileWarps * kWarp, Cfg::SMEM, s>>>(grad_rows, reso, xyz, stride, perm, cell_start, shift, log2_cells, scratch);
  T2H_CHECK_LAUNCH();
  const int64_t threads = n_seg * (C / 4);
  sample_bwd_merge_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(scratch, reso, C, log2_cells, n_seg, 1, grad_plane);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_bilinear_sample_bwd(const float* grad_rows, int64_t n_points, int reso, int C,
                                       const float* xyz_sorted, int64_t point_stride, const int32_t* perm,
                                       const int32_t* cell_start, int64_t n_seg, int shift, int morton,
                                       void* workspace, size_t workspace_bytes, float* grad_plane,
                                       t2h_stream_t stream) {
  if (!grad_rows || !xyz_sorted || !cell_start || !grad_plane || reso <= 0 || point_stride < 2 || (point_stride & 1) ||
      n_points < 0 || n_seg < 0 ||
      shift < 0 || (shift & 1) || (shift && !morton) || n_seg % ((int64_t)reso * reso))
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  {
    const int T = tiled_T(reso, C, morton, n_points, n_seg);
    static const int force_gather = []() { const char* e = getenv("T2H_SAMPLE_BWD_GATHER"); return e ? atoi(e) : 0; }();
    static const int mode = []() { const char* e = getenv("T2H_SAMPLE_BWD_MODE"); return e ? atoi(e) : 0; }();  // ablation
    if (morton && workspace && !force_gather && mode != 1 && C >= 32 && !(C & (C - 1)) && C <= 1024) {
      // cell-parallel register accumulation; block edge T with T*T*C <= 1024 floats x 9 partials of staging
      if (workspace_bytes < t2h_bilinear_sample_bwd_workspace_bytes(reso, C, n_seg, morton)) return T2H_ERR_WORKSPACE_TOO_SMALL;
      int l2c = 0;
      while ((1 << l2c) < reso) ++l2c;
      l2c *= 2;
      cudaStream_t st = (cudaStream_t)stream;
      float* scr = (float*)workspace;
#define T2H_CELL(CC, TT) \
  if (reso >= TT) return launch_cell<CC, TT>(grad_rows, reso, xyz_sorted, point_stride, perm, cell_start, n_points, n_seg, shift, l2c, scr, grad_plane, st)
      switch (C) {
        case 32: T2H_CELL(32, 4); break;
        case 64: T2H_CELL(64, 4); break;
        case 128: T2H_CELL(128, 2); break;
        case 256: T2H_CELL(256, 2); break;
        case 512: T2H_CELL(512, 1); break;
        case 1024: T2H_CELL(1024, 1); break;
        default: break;
      }
#undef T2H_CELL
    }
    // scatter tile with per-point shared-memory updates (kept for ablation: T2H_SAMPLE_BWD_MODE=1)
    if (T == 8 && workspace && !force_gather && mode == 1) {
      if (workspace_bytes < t2h_bilinear_sample_bwd_workspace_bytes(reso, C, n_seg, morton)) return T2H_ERR_WORKSPACE_TOO_SMALL;
      int l2c = 0;
      while ((1 << l2c) < reso) ++l2c;
      l2c *= 2;
      cudaStream_t st = (cudaStream_t)stream;
      float* scr = (float*)workspace;
#define T2H_TILED(CC, TT) return launch_tiled<CC, TT>(grad_rows, reso, xyz_sorted, point_stride, perm, cell_start, n_seg, shift, l2c, scr, grad_plane, st)
#define T2H_TILED_C(CC)                                             \
  case CC:                                                          \
    if (T == 8) { if constexpr (CC <= 512) T2H_TILED(CC, 8); }     \
    if (T == 4) T2H_TILED(CC, 4);                                   \
    if (T == 2) T2H_TILED(CC, 2);                                   \
    T2H_TILED(CC, 1);
      switch (C) {
        T2H_TILED_C(32)
        T2H_TILED_C(64)
        T2H_TILED_C(128)
        T2H_TILED_C(256)
        T2H_TILED_C(512)
        T2H_TILED_C(1024)
        default: break;
      }
#undef T2H_TILED_C
#undef T2H_TILED
    }
  }
  // gather fallback (row-major keys, odd channel counts): ~9 neighbour cells are scanned per plane cell;
  // split the scan over several warps on coarse levels
  const int64_t avg = n_points / n_seg;
  int wps = 1;
  while (wps < kSampleWarps && avg >= 4 * wps) wps *= 2;
  const unsigned blocks = (unsigned)((n_seg * wps + kSampleWarps - 1) / kSampleWarps);
  cudaStream_t s = (cudaStream_t)stream;
  int log2_cells = 0;
  while ((1 << log2_cells) < reso) ++log2_cells;
  log2_cells *= 2;
#define T2H_SBWD(W) sample_bwd_kernel<RS, W><<<blocks, kSampleWarps * kWarp, 0, s>>>( \
      grad_rows, reso, xyz_sorted, point_stride, perm, cell_start, n_seg, shift, morton, log2_cells, grad_plane)
  T2H_DISPATCH_ROWSHAPE(C, {
    if (wps == 1) T2H_SBWD(1); else if (wps == 2) T2H_SBWD(2); else if (wps == 4) T2H_SBWD(4); else T2H_SBWD(8);
  });
#undef T2H_SBWD
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

static inline float axis_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f; }

extern "C" int t2h_upsample_bilinear_fwd(const float* in, int B, int h, int w, int C, int out_h, int out_w,
                                         float* out, t2h_stream_t stream) {
  if (!in || !out || B < 0 || h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (B == 0) return T2H_OK;
  const float sh = axis_scale(h, out_h), sw = axis_scale(w, out_w);
  const int64_t total = (int64_t)B * out_h * out_w;
  T2H_DISPATCH_ROWSHAPE(C, {
    const int64_t per_block = 8 * RS::RPI;
    upsample_fwd_kernel<RS><<<(unsigned)((total + per_block - 1) / per_block), 256, 0, (cudaStream_t)stream>>>(
        in, B, h, w, out_h, out_w, sh, sw, out);
  });
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_upsample_bilinear_bwd(const float* grad_out, int B, int h, int w, int C, int out_h, int out_w,
                                         float* grad_in, t2h_stream_t stream) {
  if (!grad_out || !grad_in || B < 0 || h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (B == 0) return T2H_OK;
  const float sh = axis_scale(h, out_h), sw = axis_scale(w, out_w);
  // inverse scales only bound the candidate window (a degenerate axis maps every output to input 0)
  const float inv_sh = sh > 0.f ? 1.0f / sh : (float)out_h, inv_sw = sw > 0.f ? 1.0f / sw : (float)out_w;
  const int64_t total = (int64_t)B * h * w;
  T2H_DISPATCH_ROWSHAPE(C, {
    const int64_t per_block = 8 * RS::RPI;
    upsample_bwd_kernel<RS><<<(unsigned)((total + per_block - 1) / per_block), 256, 0, (cudaStream_t)stream>>>(
        grad_out, B, h, w, out_h, out_w, sh, sw, inv_sh, inv_sw, grad_in);
  });
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}
