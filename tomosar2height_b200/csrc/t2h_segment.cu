// Deterministic segmented reductions over cell-sorted points (S1 / S2 of SURVEY §2.2).
//
// Replace torch_scatter.scatter_max + gather (pointnet.py:92-99) and torch_scatter.scatter_mean
// (pointnet.py:101-111, alto.py:76-88, alto.py:187-197) and their autograd backwards.
// One warp owns one cell (segment); a row of C floats is spread over LPR lanes as float4s, so a
// warp consumes RPI = 32/LPR rows per iteration with fully coalesced 16-byte accesses.  No
// atomics: the order of every floating-point sum is fixed by the (stable) sort.
#include "t2h_common.cuh"

namespace t2h {

constexpr int kSegWarps = 8;  // warps (= segments) per CTA

struct SegGeom {
  const int32_t* perm;        // sorted position -> row (nullptr: identity)
  const int32_t* tie;         // sorted position -> original point index, for argmax ties when the
                              // sorted order inside a segment is not the point order (nullptr: it is)
  const int32_t* cell_start;  // finest-level table
  int64_t n_seg;
  int shift;   // 2k for level r = R >> k
  int morton;
  int reso;    // resolution r of THIS level
  int log2_cells;  // log2(r*r) when morton (r is a power of two): tile / cell split by shifts, no 64-bit division
};

// plane row (row-major (b, y, x)) of segment `seg` (key order of this level)
__device__ __forceinline__ int64_t plane_row(const SegGeom& g, int64_t seg) {
  if (!g.morton) return seg;
  const int64_t b = seg >> g.log2_cells;
  const uint32_t code = (uint32_t)(seg - (b << g.log2_cells));
  const int ix = (int)compact1by1(code), iy = (int)compact1by1(code >> 1);
  return (b << g.log2_cells) + (int64_t)iy * g.reso + ix;
}

// Cells with more than kHeavyMax rows (a facade in a clustered tile holds thousands of points) are not walked by
// their one warp -- that warp would run long after the rest of the grid has finished -- but deferred and
// processed by all eight warps of the CTA, partial results combined through shared memory.  The (value,
// original point index) order that decides the argmax is total, so the combination order does not matter.
constexpr int kHeavyMax = 128;

template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_max_fwd_kernel(const float* __restrict__ rows, SegGeom g, float* __restrict__ pooled,
                   float* __restrict__ plane, int32_t* __restrict__ arg) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  constexpr bool DEFER = C <= 256;  // 64 * C bytes of shared memory for the partial results
  __shared__ float hv_val[DEFER ? kSegWarps * C : 1];
  __shared__ int hv_pos[DEFER ? kSegWarps * C : 1];
  __shared__ int hv_beg[kSegWarps], hv_len[kSegWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t seg = (int64_t)blockIdx.x * kSegWarps + warp;
  const bool valid = seg < g.n_seg;
  const int sub = lane / LPR, l = lane % LPR;
  int beg = 0, end = 0;
  if (valid) { beg = g.cell_start[seg << g.shift]; end = g.cell_start[(seg + 1) << g.shift]; }
  const bool heavy = DEFER && valid && end - beg > kHeavyMax;
  if (DEFER && lane == 0) { hv_len[warp] = heavy ? end - beg : 0; hv_beg[warp] = beg; }

  float best[CH][4];
  int bpos[CH][4];
  auto reset = [&]() {
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) { best[c][k] = -FLT_MAX; bpos[c][k] = INT32_MAX; }
  };
  // does candidate (ov, op) beat (bv, bp)?  larger value; equal values -> smaller original point index
  // (= smaller sorted position when the sorted order inside the segment is the point order)
  auto beats = [&](float ov, int op, float bv, int bp) -> bool {
    if (ov > bv) return true;
    if (ov == bv && op != INT32_MAX && bp != INT32_MAX) return g.tie ? (g.tie[op] < g.tie[bp]) : (op < bp);
    return false;
  };
  auto scan = [&](int first, int last, int step) {
    for (int i = first; i < last; i += step) {
      const int64_t row = g.perm ? (int64_t)g.perm[i] : (int64_t)i;
      const float* src = rows + row * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float4 v = ld4(src + c * LPR * 4);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // strict >: the first row wins ties (rows of one fine cell are in point order); inside a
          // coarser segment equal values are resolved by the original point index
          bool take = e[k] > best[c][k];
          if (g.tie && e[k] == best[c][k] && bpos[c][k] != INT32_MAX) take = g.tie[i] < g.tie[bpos[c][k]];
          if (take) { best[c][k] = e[k]; bpos[c][k] = i; }
        }
      }
    }
  };
  auto merge_subs = [&]() {  // the RPI sub-rows of the warp
#pragma unroll
    for (int off = LPR; off < kWarp; off <<= 1) {
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float ov = __shfl_xor_sync(0xffffffffu, best[c][k], off);
          const int op = __shfl_xor_sync(0xffffffffu, bpos[c][k], off);
          if (beats(ov, op, best[c][k], bpos[c][k])) { best[c][k] = ov; bpos[c][k] = op; }
        }
    }
  };
  // plane / arg of segment `sg` (written by the sub-row 0 lanes when `writer`); best becomes the pooled value
  auto emit = [&](int64_t sg, bool empty, bool writer) {
    const int64_t prow = plane_row(g, sg);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 v;
      int4 a;
      int* ap = &a.x;
      float* vp = &v.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool none = empty || bpos[c][k] == INT32_MAX;
        vp[k] = none ? 0.0f : best[c][k];
        ap[k] = none ? -1 : (g.perm ? g.perm[bpos[c][k]] : bpos[c][k]);
        best[c][k] = vp[k];
      }
      if (writer && sub == 0) {
        const int64_t o = prow * C + (c * LPR + l) * 4;
        if (plane) st4(plane + o, v);
        *reinterpret_cast<int4*>(arg + o) = a;
      }
    }
  };
  auto gather_back = [&](int first, int last, int step) {
    for (int i = first; i < last; i += step) {
      const int64_t row = g.perm ? (int64_t)g.perm[i] : (int64_t)i;
      float* dst = pooled + row * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        st4(dst + c * LPR * 4, make_float4(best[c][0], best[c][1], best[c][2], best[c][3]));
    }
  };

  if (valid && !heavy) {
    reset();
    scan(beg + sub, end, RPI);
    merge_subs();
    emit(seg, beg >= end, true);
    if (pooled) gather_back(beg + sub, end, RPI);
  }
  if constexpr (DEFER) {
    __syncthreads();
    for (int w = 0; w < kSegWarps; ++w) {
      const int hl = hv_len[w];
      if (hl == 0) continue;  // uniform across the CTA
      const int hb = hv_beg[w];
      reset();
      scan(hb + warp * RPI + sub, hb + hl, kSegWarps * RPI);
      merge_subs();
      if (sub == 0) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            hv_val[warp * C + (c * LPR + l) * 4 + k] = best[c][k];
            hv_pos[warp * C + (c * LPR + l) * 4 + k] = bpos[c][k];
          }
      }
      __syncthreads();
      // every warp folds the eight partials (it needs the result for the gather-back of its rows)
      for (int q = 0; q < kSegWarps; ++q) {
        if (q == warp) continue;
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float ov = hv_val[q * C + (c * LPR + l) * 4 + k];
            const int op = hv_pos[q * C + (c * LPR + l) * 4 + k];
            if (beats(ov, op, best[c][k], bpos[c][k])) { best[c][k] = ov; bpos[c][k] = op; }
          }
      }
      emit((int64_t)blockIdx.x * kSegWarps + w, false, warp == 0);
      if (pooled) gather_back(hb + warp * RPI + sub, hb + hl, kSegWarps * RPI);
      __syncthreads();  // the partial buffers are free again
    }
  }
}

template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_max_bwd_kernel(const float* __restrict__ grad_pooled, const float* __restrict__ grad_plane, SegGeom g,
                   const int32_t* __restrict__ arg, float* __restrict__ grad_rows) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  constexpr bool DEFER = C <= 256;
  __shared__ float4 hv_sum[DEFER ? kSegWarps * (C / 4) : 1];
  __shared__ int hv_beg[kSegWarps], hv_len[kSegWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t seg = (int64_t)blockIdx.x * kSegWarps + warp;
  const bool valid = seg < g.n_seg;
  const int sub = lane / LPR, l = lane % LPR;
  int beg = 0, end = 0;
  if (valid) { beg = g.cell_start[seg << g.shift]; end = g.cell_start[(seg + 1) << g.shift]; }
  const bool heavy = DEFER && valid && end - beg > kHeavyMax;
  if (DEFER && lane == 0) { hv_len[warp] = heavy ? end - beg : 0; hv_beg[warp] = beg; }

  float4 acc[CH];
  auto reset = [&]() {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sum_rows = [&](int first, int last, int step) {  // pooled-gradient rows, then the RPI sub-rows of the warp
    for (int i = first; i < last; i += step) {
      const int64_t row = g.perm ? (int64_t)g.perm[i] : (int64_t)i;
      const float* src = grad_pooled + row * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float4 v = ld4(src + c * LPR * 4);
        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
      }
    }
#pragma unroll
    for (int off = LPR; off < kWarp; off <<= 1)
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float4 o = shfl_xor4(acc[c], off);
        acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
      }
  };
  // add the plane gradient, then route the total to the saved argmax row of every channel
  auto route = [&](int64_t sg, int first, int last, int step) {
    const int64_t prow = plane_row(g, sg);
    int4 a[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int64_t o = prow * C + (c * LPR + l) * 4;
      a[c] = *reinterpret_cast<const int4*>(arg + o);
      if (grad_plane) {
        float4 v = ld4(grad_plane + o);
        acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
      }
    }
    for (int i = first; i < last; i += step) {
      const int row32 = g.perm ? g.perm[i] : i;
      float* dst = grad_rows + (int64_t)row32 * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float4 v;
        v.x = a[c].x == row32 ? acc[c].x : 0.f;
        v.y = a[c].y == row32 ? acc[c].y : 0.f;
        v.z = a[c].z == row32 ? acc[c].z : 0.f;
        v.w = a[c].w == row32 ? acc[c].w : 0.f;
        st4(dst + c * LPR * 4, v);
      }
    }
  };

  if (valid && !heavy && beg < end) {
    reset();
    if (grad_pooled) sum_rows(beg + sub, end, RPI);
    route(seg, beg + sub, end, RPI);
  }
  if constexpr (DEFER) {
    __syncthreads();
    for (int w = 0; w < kSegWarps; ++w) {
      const int hl = hv_len[w];
      if (hl == 0) continue;  // uniform across the CTA
      const int hb = hv_beg[w];
      reset();
      if (grad_pooled) {
        // contiguous slice per warp, partial sums added in warp order: a fixed order, like the light path
        const int slice = (hl + kSegWarps - 1) / kSegWarps;
        const int my_beg = min(hb + warp * slice, hb + hl), my_end = min(my_beg + slice, hb + hl);
        sum_rows(my_beg + sub, my_end, RPI);
        if (sub == 0) {
#pragma unroll
          for (int c = 0; c < CH; ++c) hv_sum[warp * (C / 4) + c * LPR + l] = acc[c];
        }
        __syncthreads();
        reset();
        for (int q = 0; q < kSegWarps; ++q)
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float4 o = hv_sum[q * (C / 4) + c * LPR + l];
            acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
          }
      }
      route((int64_t)blockIdx.x * kSegWarps + w, hb + warp * RPI + sub, hb + hl, kSegWarps * RPI);
      __syncthreads();  // hv_sum is free again
    }
  }
}

constexpr int kHeavy = 64;  // rows; longer segments are processed by the whole CTA (skewed tiles)

// acc += rows[beg + sub, beg + sub + RPI, ...) ; two rows in flight per lane
template <class RS>
__device__ __forceinline__ void accum_range(const float* __restrict__ rows, const int32_t* __restrict__ perm, int beg, int end,
                                            int sub, int l, float4 (&acc)[RS::CH]) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  int i = beg + sub;
  for (; i + RPI < end; i += 2 * RPI) {
    const int64_t r0 = perm ? (int64_t)perm[i] : (int64_t)i;
    const int64_t r1 = perm ? (int64_t)perm[i + RPI] : (int64_t)(i + RPI);
    float4 v0[CH], v1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      v0[c] = ld4(rows + r0 * C + (c * LPR + l) * 4);
      v1[c] = ld4(rows + r1 * C + (c * LPR + l) * 4);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      acc[c].x += v0[c].x; acc[c].y += v0[c].y; acc[c].z += v0[c].z; acc[c].w += v0[c].w;
      acc[c].x += v1[c].x; acc[c].y += v1[c].y; acc[c].z += v1[c].z; acc[c].w += v1[c].w;
    }
  }
  for (; i < end; i += RPI) {
    const int64_t r0 = perm ? (int64_t)perm[i] : (int64_t)i;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 v = ld4(rows + r0 * C + (c * LPR + l) * 4);
      acc[c].x += v.x; acc[c].y += v.y; acc[c].z += v.z; acc[c].w += v.w;
    }
  }
}

// sum the RPI sub-rows of a warp (fixed xor tree)
template <class RS>
__device__ __forceinline__ void warp_combine(float4 (&acc)[RS::CH]) {
#pragma unroll
  for (int off = RS::LPR; off < kWarp; off <<= 1)
#pragma unroll
    for (int c = 0; c < RS::CH; ++c) {
      float4 o = shfl_xor4(acc[c], off);
      acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
    }
}

// Segment sum / mean.  WPS warps cooperate on one segment (coarse levels hold hundreds of points per
// cell): each warp reduces a contiguous slice of the segment's rows, the slices are combined through
// shared memory in slice order, so the summation order stays fixed.  With WPS == 1 (fine levels) a
// segment longer than kHeavy rows -- a facade in a clustered tile -- is deferred and reduced by all
// eight warps of the CTA the same way.
template <class RS, int WPS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_reduce_fwd_kernel(const float* __restrict__ rows, SegGeom g, int mean, float* __restrict__ plane) {
  constexpr int LPR = RS::LPR, CH = RS::CH, C = RS::C;
  constexpr int SEGS = kSegWarps / WPS;  // segments per CTA
  __shared__ float4 part_sum[kSegWarps * (C / 4)];
  __shared__ int hv_beg[kSegWarps], hv_len[kSegWarps];
  __shared__ long long hv_row[kSegWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t seg = (int64_t)blockIdx.x * SEGS + warp / WPS;
  const int part = warp % WPS;
  const bool valid = seg < g.n_seg;
  const int sub = lane / LPR, l = lane % LPR;
  int beg = 0, end = 0;
  if (valid) { beg = g.cell_start[seg << g.shift]; end = g.cell_start[(seg + 1) << g.shift]; }
  const int len = end - beg;

  auto finish = [&](float4 (&acc)[CH], int n_rows, int64_t prow) {
    const float inv = mean ? __fdiv_rn(1.0f, (float)max(n_rows, 1)) : 1.0f;  // one division, <= 1 ulp from sum / count
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 v = acc[c];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      st4(plane + prow * C + (c * LPR + l) * 4, v);
    }
  };
  // slices of `n_parts` warps -> part 0, in slice order
  auto combine_parts = [&](float4 (&acc)[CH], int first_warp, int n_parts, int my_part) {
    if (sub == 0 && my_part > 0) {
#pragma unroll
      for (int c = 0; c < CH; ++c) part_sum[warp * (C / 4) + c * LPR + l] = acc[c];
    }
    __syncthreads();
    if (my_part == 0 && sub == 0) {
      for (int q = 1; q < n_parts; ++q)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float4 o = part_sum[(first_warp + q) * (C / 4) + c * LPR + l];
          acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
        }
    }
  };

  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  // light segments: the WPS warps of the group; heavy ones (skewed tiles): deferred to the whole CTA
  const int group = warp / WPS;
  const bool heavy = WPS < kSegWarps && valid && len > kHeavy * WPS;
  if (lane == 0 && part == 0) {
    hv_len[group] = heavy ? len : 0;
    hv_beg[group] = beg;
    hv_row[group] = valid ? plane_row(g, seg) : 0;
  }
  if (valid && !heavy) {
    const int slice = (len + WPS - 1) / WPS;
    const int my_beg = min(beg + part * slice, end), my_end = min(my_beg + slice, end);
    accum_range<RS>(rows, g.perm, my_beg, my_end, sub, l, acc);
    warp_combine<RS>(acc);
  }
  if (WPS > 1) combine_parts(acc, warp - part, WPS, part);  // contains the CTA barrier
  else __syncthreads();
  if (valid && !heavy && part == 0 && sub == 0) finish(acc, len, plane_row(g, seg));
  if (WPS == kSegWarps) return;
  for (int w = 0; w < SEGS; ++w) {
    const int hl = hv_len[w];
    if (hl == 0) continue;  // uniform across the CTA
    const int hb = hv_beg[w], slice = (hl + kSegWarps - 1) / kSegWarps;
    const int my_beg = min(hb + warp * slice, hb + hl), my_end = min(my_beg + slice, hb + hl);
    __syncthreads();        // part_sum is free again
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    accum_range<RS>(rows, g.perm, my_beg, my_end, sub, l, acc);
    warp_combine<RS>(acc);
    combine_parts(acc, 0, kSegWarps, warp);
    if (warp == 0 && sub == 0) finish(acc, hl, hv_row[w]);
  }
}

template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_broadcast_kernel(const float* __restrict__ plane, SegGeom g, int mean, float* __restrict__ rows) {
  // Light cells: one warp writes the cell's rows (CTA z = 0 only).  Cells with more than kHeavyMax rows are
  // deferred: all eight warps of the CTA -- and, on coarse levels, the gridDim.y CTAs that share the cell
  // group -- write interleaved row slices, so a facade's thousands of rows do not hang on one warp.
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  __shared__ int hv_beg[kSegWarps], hv_len[kSegWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t seg = (int64_t)blockIdx.x * kSegWarps + warp;
  const bool valid = seg < g.n_seg;
  const int sub = lane / LPR, l = lane % LPR;
  int beg = 0, end = 0;
  if (valid) { beg = g.cell_start[seg << g.shift]; end = g.cell_start[(seg + 1) << g.shift]; }
  const bool heavy = valid && end - beg > kHeavyMax;
  if (lane == 0) { hv_len[warp] = heavy ? end - beg : 0; hv_beg[warp] = beg; }

  auto write_rows = [&](int64_t sg, int n_rows, int first, int last, int step) {
    const float inv = mean ? __fdiv_rn(1.0f, (float)n_rows) : 1.0f;
    const int64_t prow = plane_row(g, sg);
    float4 v[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      v[c] = ld4(plane + prow * C + (c * LPR + l) * 4);
      v[c].x *= inv; v[c].y *= inv; v[c].z *= inv; v[c].w *= inv;
    }
    for (int i = first; i < last; i += step) {
      const int64_t row = g.perm ? (int64_t)g.perm[i] : (int64_t)i;
#pragma unroll
      for (int c = 0; c < CH; ++c) st4(rows + row * C + (c * LPR + l) * 4, v[c]);
    }
  };
  if (valid && !heavy && beg < end && blockIdx.y == 0) write_rows(seg, end - beg, beg + sub, end, RPI);
  __syncthreads();
  const int slot = (int)blockIdx.y * kSegWarps + warp, n_slots = (int)gridDim.y * kSegWarps;
  for (int w = 0; w < kSegWarps; ++w) {
    const int hl = hv_len[w];
    if (hl == 0) continue;
    const int hb = hv_beg[w];
    write_rows((int64_t)blockIdx.x * kSegWarps + w, hl, hb + slot * RPI + sub, hb + hl, n_slots * RPI);
  }
}

static int check_geom(const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso) {
  if (!cell_start || n_seg < 0 || shift < 0 || shift > 30 || (shift & 1) || C <= 0 || reso <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (shift && !morton) return T2H_ERR_INVALID_ARGUMENT;  // only Morton keys nest across levels
  if (morton && ((reso & (reso - 1)) || n_seg % ((int64_t)reso * reso))) return T2H_ERR_INVALID_ARGUMENT;
  return T2H_OK;
}

static inline int log2_cells_of(int reso) {
  int l = 0;
  while ((1 << l) < reso) ++l;
  return 2 * l;
}

static inline unsigned seg_blocks(int64_t n_seg) { return (unsigned)((n_seg + kSegWarps - 1) / kSegWarps); }

// warps per segment from the mean segment length: ~8+ rows per warp, at most the whole CTA
static inline int warps_per_segment(int64_t n_rows, int64_t n_seg) {
  const int64_t avg = n_seg > 0 ? n_rows / n_seg : 0;
  int wps = 1;
  while (wps < kSegWarps && avg >= 8 * wps) wps *= 2;
  return wps;
}

}  // namespace t2h

using namespace t2h;

extern "C" int t2h_seg_max_fwd(const float* rows, const int32_t* perm, const int32_t* tie_rank,
                               const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                               float* pooled, float* plane, int32_t* arg, t2h_stream_t stream) {
  int st = check_geom(cell_start, n_seg, shift, C, morton, reso);
  if (st) return st;
  if (!rows || !arg) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  SegGeom g{perm, tie_rank, cell_start, n_seg, shift, morton, reso, log2_cells_of(reso)};
  T2H_DISPATCH_ROWSHAPE(C, seg_max_fwd_kernel<RS><<<seg_blocks(n_seg), kSegWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               rows, g, pooled, plane, arg));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_max_bwd(const float* grad_pooled, const float* grad_plane, const int32_t* perm,
                               const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                               const int32_t* arg, float* grad_rows, t2h_stream_t stream) {
  int st = check_geom(cell_start, n_seg, shift, C, morton, reso);
  if (st) return st;
  if (!arg || !grad_rows || (!grad_pooled && !grad_plane)) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  SegGeom g{perm, nullptr, cell_start, n_seg, shift, morton, reso, log2_cells_of(reso)};
  T2H_DISPATCH_ROWSHAPE(C, seg_max_bwd_kernel<RS><<<seg_blocks(n_seg), kSegWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               grad_pooled, grad_plane, g, arg, grad_rows));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_reduce_fwd(const float* rows, int64_t n_rows, const int32_t* perm, const int32_t* cell_start,
                                  int64_t n_seg, int shift, int C, int morton, int reso, int mean, float* plane,
                                  t2h_stream_t stream) {
  int st = check_geom(cell_start, n_seg, shift, C, morton, reso);
  if (st) return st;
  if (!rows || !plane) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  SegGeom g{perm, nullptr, cell_start, n_seg, shift, morton, reso, log2_cells_of(reso)};
  const int wps = warps_per_segment(n_rows, n_seg);
  const unsigned blocks = (unsigned)((n_seg * wps + kSegWarps - 1) / kSegWarps);
  cudaStream_t s = (cudaStream_t)stream;
  T2H_DISPATCH_ROWSHAPE(C, {
    if (wps == 1) seg_reduce_fwd_kernel<RS, 1><<<blocks, kSegWarps * kWarp, 0, s>>>(rows, g, mean, plane);
    else if (wps == 2) seg_reduce_fwd_kernel<RS, 2><<<blocks, kSegWarps * kWarp, 0, s>>>(rows, g, mean, plane);
    else if (wps == 4) seg_reduce_fwd_kernel<RS, 4><<<blocks, kSegWarps * kWarp, 0, s>>>(rows, g, mean, plane);
    else seg_reduce_fwd_kernel<RS, 8><<<blocks, kSegWarps * kWarp, 0, s>>>(rows, g, mean, plane);
  });
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_broadcast(const float* plane, const int32_t* perm, const int32_t* cell_start, int64_t n_seg,
                                 int shift, int C, int morton, int reso, int mean, float* rows,
                                 t2h_stream_t stream) {
  int st = check_geom(cell_start, n_seg, shift, C, morton, reso);
  if (st) return st;
  if (!rows || !plane) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  SegGeom g{perm, nullptr, cell_start, n_seg, shift, morton, reso, log2_cells_of(reso)};
  // coarse levels (few cells, many rows each): four CTAs share every group of cells for its heavy ones
  const dim3 grid(seg_blocks(n_seg), n_seg < 16384 ? 4 : 1);
  T2H_DISPATCH_ROWSHAPE(C, seg_broadcast_kernel<RS><<<grid, kSegWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               plane, g, mean, rows));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}
