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

constexpr int kTileWarps = 8;

// ---- G2, fine levels (few rows per cell): row-balanced walk with warp-private shared-memory tiles --------
// The cells of a level are grouped into Morton blocks of TW x TW cells; the points of a block are ONE contiguous
// range of the sorted order and all their taps fall into the block's haloed (TW+2) x (TW+2) tile (taps of a point
// of cell cx lie in columns cx-1..cx+1).  As everywhere else the work is split over the ROWS: warp w of the grid
// owns the sorted positions [w * kRowChunk, (w + 1) * kRowChunk) and walks the blocks they belong to, one after the
// other, accumulating each into a private tile in shared memory.  A row of C floats is LPR = C/4 lanes x float4;
// the G = 32/LPR lane groups of the warp all load the SAME row and each adds it into a different one of the row's
// four taps (G = 4: one tap each, G = 1: four taps in turn), so no two lanes ever touch one accumulator at the
// same time and every accumulator sees its contributions in sorted point order: atomic-free, deterministic.
// A block that lies inside the chunk is complete -> its tile goes to blocks[block]; a block that crosses a chunk
// border (a facade end holds thousands of rows in a few cells) leaves a partial tile per chunk in slots[chunk][s]
// (s = 0: the block entered from the left, s = 1: it starts here and leaves to the right).  Warps are independent
// work items -- no CTA-wide barrier, so a fat block never holds an SM hostage.
// sample_bwd_wtile_merge_kernel then builds every plane cell from the (<= 4) block tiles that cover it, adding the
// chunk partials of a crossing block in chunk order.
constexpr int kRowWarps = 4;    // warps per CTA of the tile walk; every warp is an independent work item
constexpr int kRowChunk = 128;  // sorted positions per warp

template <int C, int TW>
struct WTile {
  static constexpr int LPR = C / 4, G = 32 / LPR, TAPS = 4 / G;
  static constexpr int TWH = TW + 2;
  static constexpr int TILE_FLOATS = TWH * TWH * C;
  static constexpr int LOG2_BLOCK = (TW == 4) ? 4 : 2;
  static constexpr int STAGE_WORDS = 32 * 9;              // per warp: 32 rows x (4 tap cells, 4 weights, row)
  static constexpr int WARP_BYTES = (TILE_FLOATS + STAGE_WORDS) * 4;   // private tile + staging
  static constexpr int SMEM = kRowWarps * WARP_BYTES;
};

// rows [r0, r1) of the sorted order (all inside the block whose haloed tile origin is (bx0 - 1, by0 - 1)) -> tile.
// Lane j resolves the taps of row j of a 32-row batch once and parks them in the warp's staging arrays; the row
// loads of the NEXT U rows are in flight while the current U rows are added (double buffer).
template <int C, int TW>
__device__ __forceinline__ void wtile_accumulate(float* __restrict__ tile, int* __restrict__ st_cell, float* __restrict__ st_w,
                                                 int* __restrict__ st_row, const float* __restrict__ grad_rows, int reso,
                                                 const float* __restrict__ xyz, int64_t stride, const int32_t* __restrict__ perm,
                                                 int r0, int r1, int bx0, int by0, int lane) {
  using Cfg = WTile<C, TW>;
  constexpr int LPR = Cfg::LPR, TAPS = Cfg::TAPS, TWH = Cfg::TWH;
  constexpr int U = 4;
  const int sub = lane / LPR, l = lane % LPR;
  const float* gbase = grad_rows + l * 4;
  for (int base_i = r0; base_i < r1; base_i += kWarp) {
    const int nb = min(kWarp, r1 - base_i);
    if (lane < nb) {
      const int i = base_i + lane;
      const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + (int64_t)i * stride));
      const Taps t = make_taps(pxy.x, pxy.y, reso);
      const bool in_x = t.x0 + 1 < reso, in_y = t.y0 + 1 < reso;
      // local coordinates in the haloed tile; clamp defensively (a point is always within one cell of its own)
      const int lx = min(max(t.x0 - bx0 + 1, 0), TW), ly = min(max(t.y0 - by0 + 1, 0), TW);
      const int c0 = (ly * TWH + lx) * C;
      // taps beyond the last row / column do not exist (ATen skips them): cell -1
      *reinterpret_cast<int4*>(st_cell + lane * 4) =
          make_int4(c0, in_x ? c0 + C : -1, in_y ? c0 + TWH * C : -1, (in_x && in_y) ? c0 + TWH * C + C : -1);
      *reinterpret_cast<float4*>(st_w + lane * 4) = make_float4(__fmul_rn(t.wx0, t.wy0), __fmul_rn(t.wx1, t.wy0),
                                                                __fmul_rn(t.wx0, t.wy1), __fmul_rn(t.wx1, t.wy1));
      st_row[lane] = perm ? __ldg(perm + i) : i;
    }
    __syncwarp();
    float4 buf_a[U], buf_b[U];
    auto load = [&](float4 (&dst)[U], int j0) {
#pragma unroll
      for (int u = 0; u < U; ++u) dst[u] = ld4_stream(gbase + (int64_t)st_row[min(j0 + u, nb - 1)] * C);
    };
    auto add = [&](const float4 (&src)[U], int j0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u < nb) {
#pragma unroll
          for (int tt = 0; tt < TAPS; ++tt) {
            const int t = sub * TAPS + tt;
            const int cell = st_cell[(j0 + u) * 4 + t];
            const float w = st_w[(j0 + u) * 4 + t];
            if (cell >= 0) {
              float4* a = reinterpret_cast<float4*>(tile + cell + l * 4);
              float4 v = *a;
              v.x += w * src[u].x; v.y += w * src[u].y; v.z += w * src[u].z; v.w += w * src[u].w;
              *a = v;
            }
          }
        }
        __syncwarp();  // the next row may hit the accumulators another lane group just wrote
      }
    };
    load(buf_a, 0);
    for (int j0 = 0; j0 < nb; j0 += 2 * U) {  // ping-pong: the loads of the next U rows fly while U rows are added
      if (j0 + U < nb) load(buf_b, j0 + U);
      add(buf_a, j0);
      if (j0 + 2 * U < nb) load(buf_a, j0 + 2 * U);
      if (j0 + U < nb) add(buf_b, j0 + U);
    }
    __syncwarp();      // the staging arrays are rewritten by the next batch
  }
}

template <int C, int TW>
__global__ void __launch_bounds__(kRowWarps * kWarp)
sample_bwd_wtile_rows_kernel(const float* __restrict__ grad_rows, int64_t n_rows, int reso, const float* __restrict__ xyz,
                             int64_t stride, const int32_t* __restrict__ perm, const int32_t* __restrict__ keys,
                             const int32_t* __restrict__ cell_start, int shift, int log2_cells,
                             float* __restrict__ blocks, float* __restrict__ slots) {
  using Cfg = WTile<C, TW>;
  extern __shared__ float wt_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = wt_smem + warp * (Cfg::WARP_BYTES / 4);
  int* st_cell = reinterpret_cast<int*>(tile + Cfg::TILE_FLOATS);  // [32][4]
  float* st_w = reinterpret_cast<float*>(st_cell + 32 * 4);        // [32][4]
  int* st_row = st_cell + 32 * 8;                                  // [32]
  const int64_t chunk = (int64_t)blockIdx.x * kRowWarps + warp;
  const int64_t first64 = chunk * kRowChunk;
  if (first64 >= n_rows) return;
  const int first = (int)first64, last = (int)min(first64 + (int64_t)kRowChunk, n_rows);
  const int bshift = shift + Cfg::LOG2_BLOCK;  // sort key -> block of this level
  int row = first;
  while (row < last) {
    const int blk = __ldg(keys + row) >> bshift;
    const int64_t key0 = (int64_t)blk << Cfg::LOG2_BLOCK;  // first cell of the block (level key incl. the image)
    const int p0 = __ldg(cell_start + (key0 << shift)), p1 = __ldg(cell_start + ((key0 + (1 << Cfg::LOG2_BLOCK)) << shift));
    const int r1 = min(p1, last);
    const uint32_t code = (uint32_t)(key0 & ((1ll << log2_cells) - 1));
    const int bx0 = (int)compact1by1(code), by0 = (int)compact1by1(code >> 1);
    for (int f = lane; f < Cfg::TILE_FLOATS / 4; f += kWarp) reinterpret_cast<float4*>(tile)[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    wtile_accumulate<C, TW>(tile, st_cell, st_w, st_row, grad_rows, reso, xyz, stride, perm, row, r1, bx0, by0, lane);
    const bool complete = p0 >= first && p1 <= last;
    float* dst = complete ? blocks + (int64_t)blk * Cfg::TILE_FLOATS
                          : slots + (chunk * 2 + (p0 < first ? 0 : 1)) * (int64_t)Cfg::TILE_FLOATS;
    for (int f = lane; f < Cfg::TILE_FLOATS / 4; f += kWarp) st4(dst + f * 4, reinterpret_cast<const float4*>(tile)[f]);
    __syncwarp();
    row = r1;
  }
}

// grad_plane[b, y, x, :] = sum over the (<= 4) blocks whose haloed tile covers (x, y), in block order (row, then
// column): an empty block contributes nothing, a block inside one chunk its tile, a block that crosses chunk borders
// the chunk partials in chunk order.
template <int C, int TW>
__global__ void __launch_bounds__(256)
sample_bwd_wtile_merge_kernel(const float* __restrict__ blocks, const float* __restrict__ slots,
                              const int32_t* __restrict__ cell_start, int shift, int log2_reso, int64_t n_cells,
                              float* __restrict__ grad_plane) {
  constexpr int TWH = TW + 2;
  constexpr int LOG2_TW = (TW == 4) ? 2 : 1, LOG2_BLOCK = 2 * LOG2_TW;
  constexpr int TILE = TWH * TWH * C;
  constexpr int C4 = C / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * C4) return;
  const int64_t cellid = gid / C4;
  const int ch = (int)(gid % C4) * 4;
  const int log2_cells = 2 * log2_reso, reso = 1 << log2_reso;
  const int64_t img = cellid >> log2_cells;
  const int rem = (int)(cellid & ((1ll << log2_cells) - 1));
  const int y = rem >> log2_reso, x = rem & (reso - 1);
  const int bx = x >> LOG2_TW, by = y >> LOG2_TW, ox = x & (TW - 1), oy = y & (TW - 1), nb = reso >> LOG2_TW;
  // the own block always covers the cell; a neighbour only when the cell sits on the facing edge (one-cell halo)
  const int qx_lo = (ox == 0 && bx > 0) ? bx - 1 : bx, qx_hi = (ox == TW - 1 && bx + 1 < nb) ? bx + 1 : bx;
  const int qy_lo = (oy == 0 && by > 0) ? by - 1 : by, qy_hi = (oy == TW - 1 && by + 1 < nb) ? by + 1 : by;
  // up to 2 x 2 candidate blocks in (row, column) order; their row ranges first, then their tiles, all in flight together
  int64_t blk[4];
  int p0[4], p1[4], off[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int qy = (k >> 1) ? qy_hi : qy_lo, qx = (k & 1) ? qx_hi : qx_lo;
    const bool dup = ((k >> 1) && qy_hi == qy_lo) || ((k & 1) && qx_hi == qx_lo);
    blk[k] = (img << (log2_cells - LOG2_BLOCK)) + (part1by1((uint32_t)qx) | (part1by1((uint32_t)qy) << 1));
    off[k] = ((y - (qy << LOG2_TW) + 1) * TWH + (x - (qx << LOG2_TW) + 1)) * C + ch;
    const int64_t key0 = blk[k] << LOG2_BLOCK;
    p0[k] = dup ? 0 : __ldg(cell_start + (key0 << shift));
    p1[k] = dup ? 0 : __ldg(cell_start + ((key0 + (1 << LOG2_BLOCK)) << shift));
  }
  float4 v[4];
  bool crossing = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p1[k] > p0[k]) {
      if (p0[k] / kRowChunk == (p1[k] - 1) / kRowChunk) v[k] = ld4(blocks + blk[k] * TILE + off[k]);
      else crossing = true;
    }
  }
  if (crossing) {  // a block that spans several chunks: its chunk partials in chunk order
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      if (p1[k] <= p0[k]) continue;
      const int c0 = p0[k] / kRowChunk, c1 = (p1[k] - 1) / kRowChunk;
      if (c0 == c1) continue;
      float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int ck = c0; ck <= c1; ck += 4) {
        float4 o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cq = min(ck + q, c1);
          o[q] = ld4(slots + ((int64_t)cq * 2 + ((cq * kRowChunk > p0[k]) ? 0 : 1)) * TILE + off[k]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (ck + q <= c1) { part.x += o[q].x; part.y += o[q].y; part.z += o[q].z; part.w += o[q].w; }
      }
      v[k] = part;
    }
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
  st4(grad_plane + cellid * C + ch, acc);
}

// ---- G2, coarse levels (many rows per cell): row-balanced nine-partial walk -----------------------------
// All points of one own cell (cx, cy) contribute to the same 3x3 target cells, with separable weights, so the
// backward is a SEGMENTED SUM of weighted rows with nine outputs per cell.  As in t2h_segment.cu the work is
// split over the rows: every group of LPR lanes walks kRows9 consecutive sorted positions of one channel
// group (<= 128 channels: nine float4 accumulators per lane), finishes the cells that lie inside its chunk
// (-> nine[cell][9][C]) and leaves the partials of the cells that cross a chunk border in scratch slots;
// sample_bwd_nine_fix_kernel adds those in chunk order and zero-fills empty cells; sample_bwd_nine_gather_kernel
// assembles grad_plane[t] = sum_d nine[t - d][d].  A facade with 20 000 points in one cell is walked by ~80
// independent warps; no atomics, fixed summation order.
template <int LPR> struct Rows9 { static constexpr int ROWS = (LPR == 32) ? 256 : 128; };

template <int LPR>
__global__ void __launch_bounds__(kTileWarps * kWarp)
sample_bwd_nine_rows_kernel(const float* __restrict__ grad_rows, int64_t n_rows, int C, int reso,
                            const float* __restrict__ xyz, int64_t stride, const int32_t* __restrict__ perm,
                            const int32_t* __restrict__ keys, int shift, int morton, int log2_cells,
                            float* __restrict__ nine, float* __restrict__ slots) {
  constexpr int RPI = 32 / LPR, ROWS = Rows9<LPR>::ROWS, U = 8;
  static_assert(LPR % U == 0, "a batch of LPR rows is consumed U rows at a time");
  // per warp: the resolved rows of the current batch -- {key, row, class, -} and the four tap weights
  __shared__ int4 st_meta[kTileWarps][32];
  __shared__ float4 st_w[kTileWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * kTileWarps + warp) * RPI + sub;
  // every lane group walks its own chunk; groups past the end idle through the batch loop
  const int64_t first = min(chunk * ROWS, n_rows);
  const int64_t last = min(first + (int64_t)ROWS, n_rows);
  const bool live = first < last;
  const int ch0 = blockIdx.y * (LPR * 4) + l * 4;  // this lane's four channels
  const int4* meta = &st_meta[warp][sub * LPR];
  const float4* wts = &st_w[warp][sub * LPR];

  auto key_of = [&](int64_t i) { return __ldg(keys + i) >> shift; };
  int cur = live ? key_of(first) : 0;
  const bool cont_in = live && first > 0 && key_of(first - 1) == cur;
  bool is_first = true;
  float4 acc[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&](int slot) {
    float* dst = slot < 0 ? nine + (int64_t)cur * 9 * C : slots + (chunk * 2 + slot) * 9 * (int64_t)C;
#pragma unroll
    for (int d = 0; d < 9; ++d) st4(dst + d * C + ch0, acc[d]);
  };
  auto fma4 = [](float4& a, float w, const float4& g) { a.x += w * g.x; a.y += w * g.y; a.z += w * g.z; a.w += w * g.w; };

  // batches of LPR rows: lane l of the group resolves key, row index, tap class and the four tap weights of row
  // base + l ONCE (the coordinate arithmetic is ~40 instructions) and parks them in shared memory; the group then
  // consumes the batch row by row with two broadcast loads per row
  for (int64_t base = first; __any_sync(0xffffffffu, base < last); base += LPR) {
    const int nb = (int)max((int64_t)0, min((int64_t)LPR, last - base));
    __syncwarp();  // the previous batch has been consumed
    if (l < nb) {
      const int64_t p = base + l;
      const int key = key_of(p);
      const int row = perm ? __ldg(perm + p) : (int)p;
      const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + p * stride));
      const Taps t = make_taps(pxy.x, pxy.y, reso);
      const float wx1 = (t.x0 + 1 < reso) ? t.wx1 : 0.f, wy1 = (t.y0 + 1 < reso) ? t.wy1 : 0.f;
      int cx, cy;
      cell_decode((uint32_t)key & ((1u << log2_cells) - 1u), reso, morton, cx, cy);
      // the tap origin is the own cell or the one before it, per axis
      st_meta[warp][lane] = make_int4(key, row, (t.x0 < cx ? 1 : 0) | (t.y0 < cy ? 2 : 0), 0);
      st_w[warp][lane] = make_float4(__fmul_rn(t.wx0, t.wy0), __fmul_rn(wx1, t.wy0), __fmul_rn(t.wx0, wy1), __fmul_rn(wx1, wy1));
    }
    __syncwarp();
    for (int j0 = 0; j0 < nb; j0 += U) {
      float4 gq[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j0 + u < nb) gq[u] = ld4_stream(grad_rows + (int64_t)meta[j0 + u].y * C + ch0);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u >= nb) break;
        const int4 m = meta[j0 + u];
        const float4 w = wts[j0 + u];
        if (m.x != cur) {
          flush(is_first && cont_in ? 0 : -1);
          is_first = false;
          cur = m.x;
#pragma unroll
          for (int d = 0; d < 9; ++d) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // four cases, four live partials each (acc index = 3 * (target row - cy + 1) + (target column - cx + 1))
        switch (m.z) {
          case 3:  fma4(acc[0], w.x, gq[u]); fma4(acc[1], w.y, gq[u]); fma4(acc[3], w.z, gq[u]); fma4(acc[4], w.w, gq[u]); break;
          case 2:  fma4(acc[1], w.x, gq[u]); fma4(acc[2], w.y, gq[u]); fma4(acc[4], w.z, gq[u]); fma4(acc[5], w.w, gq[u]); break;
          case 1:  fma4(acc[3], w.x, gq[u]); fma4(acc[4], w.y, gq[u]); fma4(acc[6], w.z, gq[u]); fma4(acc[7], w.w, gq[u]); break;
          default: fma4(acc[4], w.x, gq[u]); fma4(acc[5], w.y, gq[u]); fma4(acc[7], w.z, gq[u]); fma4(acc[8], w.w, gq[u]); break;
        }
      }
    }
  }
  if (!live) return;
  const bool cont_out = last < n_rows && key_of(last) == cur;
  flush(is_first && cont_in ? 0 : (cont_out ? 1 : -1));
}

// one thread per (cell, float4 of its 9 x C partials): empty -> zero; crossing a chunk border -> sum of the chunk
// partials in chunk order, four loads in flight
template <int ROWS>
__global__ void __launch_bounds__(256)
sample_bwd_nine_fix_kernel(const int32_t* __restrict__ cell_start, int64_t n_seg, int shift, int C,
                           const float* __restrict__ slots, float* __restrict__ nine) {
  const int per_cell = 9 * C / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_seg * per_cell) return;
  const int64_t seg = gid / per_cell;
  const int e = (int)(gid - seg * per_cell);
  const int beg = __ldg(cell_start + (seg << shift)), end = __ldg(cell_start + ((seg + 1) << shift));
  const int c0 = beg / ROWS, c1 = (end - 1) / ROWS;
  if (beg < end && c0 == c1) return;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  if (beg < end)
    for (int ck = c0; ck <= c1; ck += 4) {
      float4 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cq = min(ck + q, c1);
        o[q] = ld4(slots + ((int64_t)cq * 2 + ((cq * ROWS > beg) ? 0 : 1)) * 9 * C + e * 4);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (ck + q <= c1) { sum.x += o[q].x; sum.y += o[q].y; sum.z += o[q].z; sum.w += o[q].w; }
    }
  st4(nine + seg * 9 * (int64_t)C + e * 4, sum);
}

// grad_plane[b, ty, tx, :] = sum over (dy, dx) of nine[cell (tx - dx, ty - dy)][(dy + 1) * 3 + dx + 1]
__global__ void __launch_bounds__(256)
sample_bwd_nine_gather_kernel(const float* __restrict__ nine, int reso, int C, int morton, int log2_cells, int64_t n_cells,
                              float* __restrict__ grad_plane) {
  const int c4 = C / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cells * c4) return;
  const int64_t cellid = gid / c4;
  const int ch = (int)(gid - cellid * c4) * 4;
  const int64_t cells = (int64_t)reso * reso;
  const int64_t img = cellid / cells;
  const int rem = (int)(cellid - img * cells);
  const int ty = rem / reso, tx = rem - ty * reso;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int ox = tx - dx, oy = ty - dy;
      if (ox < 0 || ox >= reso || oy < 0 || oy >= reso) continue;
      const int64_t key = img * cells + cell_code((uint32_t)ox, (uint32_t)oy, reso, morton);
      const float4 v = ld4(nine + (key * 9 + (dy + 1) * 3 + (dx + 1)) * (int64_t)C + ch);
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

// Which G2 kernel serves a level.  Warp-private tiles need Morton keys, C in {32, 64, 128} and few rows per
// cell (a warp walks a whole block of cells); the nine-partial walk needs Morton keys and C = 32, 64 or a
// multiple of 128; everything else (row-major keys, C < 32) takes the 3x3 gather.
enum { G2_GATHER = 0, G2_WTILE = 1, G2_NINE = 2 };
static inline int g2_mode(int reso, int C, int morton, int64_t n_points, int64_t n_seg) {
  if (!morton || n_seg <= 0) return G2_GATHER;
  const bool nine_ok = C == 32 || C == 64 || (C % 128 == 0 && C <= 1024);
  const int64_t avg = n_points / n_seg;
  const int tw = C == 128 ? 2 : 4;
  if ((C == 32 || C == 64 || C == 128) && avg < 32 && reso >= tw) return G2_WTILE;
  return nine_ok ? G2_NINE : G2_GATHER;
}
static inline int nine_rows(int C) { return C >= 128 ? 256 : 128; }
static inline size_t align256s(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t t2h_bilinear_sample_bwd_workspace_bytes(int reso, int C, int64_t n_points, int64_t n_seg, int morton) {
  switch (g2_mode(reso, C, morton, n_points, n_seg)) {
    case G2_WTILE: {
      const int tw = C == 128 ? 2 : 4;
      const size_t tile = (size_t)(tw + 2) * (tw + 2) * C * sizeof(float);
      const int64_t chunks = (n_points + kRowChunk - 1) / kRowChunk;
      return align256s((size_t)(n_seg / (tw * tw)) * tile) + (size_t)chunks * 2 * tile + 256;
    }
    case G2_NINE: {
      const int64_t chunks = (n_points + nine_rows(C) - 1) / nine_rows(C);
      return align256s((size_t)n_seg * 9 * C * sizeof(float)) + (size_t)chunks * 2 * 9 * C * sizeof(float) + 256;
    }
    default: return 256;
  }
}

template <int C, int TW>
static int launch_wtile(const float* grad_rows, int64_t n_points, int reso, const float* xyz, int64_t stride,
                        const int32_t* perm, const int32_t* keys, const int32_t* cell_start, int64_t n_seg, int shift,
                        int log2_cells, float* scratch, float* grad_plane, cudaStream_t s) {
  using Cfg = WTile<C, TW>;
  auto kern = sample_bwd_wtile_rows_kernel<C, TW>;
  // per device and idempotent; set on every launch so that the entry point keeps no state
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM) != cudaSuccess) {
    (void)cudaGetLastError();
    return T2H_ERR_CUDA;
  }
  float* blocks = scratch;
  float* slots = (float*)((char*)scratch + align256s((size_t)(n_seg >> Cfg::LOG2_BLOCK) * Cfg::TILE_FLOATS * sizeof(float)));
  const int64_t chunks = (n_points + kRowChunk - 1) / kRowChunk;
  if (chunks > 0) {
    kern<<<(unsigned)((chunks + kRowWarps - 1) / kRowWarps), kRowWarps * kWarp, Cfg::SMEM, s>>>(
        grad_rows, n_points, reso, xyz, stride, perm, keys, cell_start, shift, log2_cells, blocks, slots);
    T2H_CHECK_LAUNCH();
  }
  const int64_t threads = n_seg * (C / 4);
  sample_bwd_wtile_merge_kernel<C, TW><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(blocks, slots, cell_start, shift,
                                                                                        log2_cells / 2, n_seg, grad_plane);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

template <int LPR>
static int launch_nine(const float* grad_rows, int64_t n_points, int reso, int C, const float* xyz, int64_t stride,
                       const int32_t* perm, const int32_t* keys, const int32_t* cell_start, int64_t n_seg, int shift,
                       int morton, int log2_cells, float* nine, float* slots, float* grad_plane, cudaStream_t s) {
  constexpr int ROWS = Rows9<LPR>::ROWS, RPI = 32 / LPR;
  const int64_t chunks = (n_points + ROWS - 1) / ROWS;
  if (chunks > 0) {
    const int64_t warps = (chunks + RPI - 1) / RPI;
    const dim3 grid((unsigned)((warps + kTileWarps - 1) / kTileWarps), (unsigned)(C / (LPR * 4)));
    sample_bwd_nine_rows_kernel<LPR><<<grid, kTileWarps * kWarp, 0, s>>>(grad_rows, n_points, C, reso, xyz, stride, perm, keys,
                                                                        shift, morton, log2_cells, nine, slots);
    T2H_CHECK_LAUNCH();
  }
  const int64_t fix_threads = n_seg * (9 * C / 4);
  sample_bwd_nine_fix_kernel<ROWS><<<(unsigned)((fix_threads + 255) / 256), 256, 0, s>>>(cell_start, n_seg, shift, C, slots, nine);
  T2H_CHECK_LAUNCH();
  const int64_t threads = n_seg * (C / 4);
  sample_bwd_nine_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(nine, reso, C, morton, log2_cells, n_seg, grad_plane);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_bilinear_sample_bwd(const float* grad_rows, int64_t n_points, int reso, int C,
                                       const float* xyz_sorted, int64_t point_stride, const int32_t* perm,
                                       const int32_t* row_keys, const int32_t* cell_start, int64_t n_seg, int shift,
                                       int morton, void* workspace, size_t workspace_bytes, float* grad_plane,
                                       t2h_stream_t stream) {
  if (!cell_start || !grad_plane || reso <= 0 || point_stride < 2 || (point_stride & 1) || n_points < 0 || n_seg < 0 ||
      (n_points > 0 && (!grad_rows || !xyz_sorted)) || n_points > INT32_MAX ||
      shift < 0 || (shift & 1) || (shift && !morton) || n_seg % ((int64_t)reso * reso))
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int log2_cells = 0;
  while ((1 << log2_cells) < reso) ++log2_cells;
  log2_cells *= 2;
  const int mode = (workspace && row_keys) ? g2_mode(reso, C, morton, n_points, n_seg) : G2_GATHER;
  if (mode != G2_GATHER && workspace_bytes < t2h_bilinear_sample_bwd_workspace_bytes(reso, C, n_points, n_seg, morton))
    return T2H_ERR_WORKSPACE_TOO_SMALL;
  if (mode == G2_WTILE) {
    float* scr = (float*)workspace;
    if (C == 32) return launch_wtile<32, 4>(grad_rows, n_points, reso, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, log2_cells, scr, grad_plane, s);
    if (C == 64) return launch_wtile<64, 4>(grad_rows, n_points, reso, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, log2_cells, scr, grad_plane, s);
    return launch_wtile<128, 2>(grad_rows, n_points, reso, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, log2_cells, scr, grad_plane, s);
  }
  if (mode == G2_NINE) {
    float* nine = (float*)workspace;
    float* slots = (float*)((char*)workspace + align256s((size_t)n_seg * 9 * C * sizeof(float)));
    if (C == 32) return launch_nine<8>(grad_rows, n_points, reso, C, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, morton, log2_cells, nine, slots, grad_plane, s);
    if (C == 64) return launch_nine<16>(grad_rows, n_points, reso, C, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, morton, log2_cells, nine, slots, grad_plane, s);
    return launch_nine<32>(grad_rows, n_points, reso, C, xyz_sorted, point_stride, perm, row_keys, cell_start, n_seg, shift, morton, log2_cells, nine, slots, grad_plane, s);
  }
  // gather fallback (row-major keys, odd channel counts): ~9 neighbour cells are scanned per plane cell;
  // split the scan over several warps on coarse levels
  const int64_t avg = n_points / n_seg;
  int wps = 1;
  while (wps < kSampleWarps && avg >= 4 * wps) wps *= 2;
  const unsigned blocks = (unsigned)((n_seg * wps + kSampleWarps - 1) / kSampleWarps);
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
