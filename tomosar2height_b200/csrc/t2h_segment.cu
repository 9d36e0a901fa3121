// Deterministic, ROW-BALANCED segmented reductions over cell-sorted points (S1 / S2 of SURVEY §2.2).
//
// Replace torch_scatter.scatter_max + gather (pointnet.py:92-99) and torch_scatter.scatter_mean
// (pointnet.py:101-111, alto.py:76-88, alto.py:187-197) and their autograd backwards.
//
// Work is split over the ROWS, not over the cells: every group of LPR lanes (a row of C floats = LPR
// lanes x CH float4) walks its own chunk of kChunk consecutive sorted positions, so a facade cell with
// thousands of points costs exactly as much per row as a cell with one.  A cell that lies completely
// inside a chunk is finished by its walker and written straight to the plane.  A cell that crosses a
// chunk border leaves one partial per chunk in a scratch slot (slot 0: the cell entered the chunk from
// the left, slot 1: it starts here and leaves to the right); a second "fix-up" launch visits every cell,
// adds the partials of the crossing ones in chunk order and zero-fills the empty ones.  No atomics: the
// order of every floating-point sum is fixed by the (stable) sort and the chunk grid.
// The backward passes and the gather-back are pure row-parallel maps (one plane row looked up per point).
#include "t2h_common.cuh"

namespace t2h {

constexpr int kSegWarps = 8;  // warps per CTA
constexpr int kChunk = 32;    // sorted positions per walker (sub-group) chunk

struct SegGeom {
  const int32_t* perm;        // sorted position -> row (nullptr: identity)
  const int32_t* tie;         // sorted position -> original point index, for argmax ties when the
                              // sorted order inside a segment is not the point order (nullptr: it is)
  const int32_t* keys;        // sorted position -> finest-level sort key
  const int32_t* cell_start;  // finest-level table
  int64_t n_rows;
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

__device__ __forceinline__ int level_key(const SegGeom& g, int64_t i) { return __ldg(g.keys + i) >> g.shift; }

// ---- S2 forward: segment sum / mean -----------------------------------------------------------------
// scratch: [n_chunks][2][C] floats (raw sums of the border cells of every chunk)
template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_reduce_rows_kernel(const float* __restrict__ rows, SegGeom g, int mean, float* __restrict__ plane,
                       float* __restrict__ scratch) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  constexpr int U = CH >= 4 ? 2 : 4;  // rows in flight per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * kSegWarps + warp) * RPI + sub;
  const int64_t first = chunk * kChunk;
  if (first >= g.n_rows) return;
  const int64_t last = min(first + (int64_t)kChunk, g.n_rows);

  int cur = level_key(g, first);
  const bool cont_in = first > 0 && level_key(g, first - 1) == cur;  // the first cell entered from the left
  bool is_first = true;
  int cnt = 0;
  float4 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto flush = [&](int slot) {  // slot < 0: the cell is complete -> plane
    if (slot < 0) {
      const float inv = mean ? __fdiv_rn(1.0f, (float)cnt) : 1.0f;  // one division, <= 1 ulp from sum / count
      float* dst = plane + plane_row(g, cur) * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        st4(dst + c * LPR * 4, make_float4(acc[c].x * inv, acc[c].y * inv, acc[c].z * inv, acc[c].w * inv));
    } else {
      float* dst = scratch + (chunk * 2 + slot) * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) st4(dst + c * LPR * 4, acc[c]);
    }
  };

  for (int64_t i = first; i < last; i += U) {
    int k[U];
    float4 v[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = min(i + u, last - 1);
      k[u] = level_key(g, p);
      const int64_t row = g.perm ? (int64_t)__ldg(g.perm + p) : p;
#pragma unroll
      for (int c = 0; c < CH; ++c) v[u][c] = ld4(rows + row * C + (c * LPR + l) * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u >= last) break;
      if (k[u] != cur) {
        flush(is_first && cont_in ? 0 : -1);
        is_first = false;
        cur = k[u];
        cnt = 0;
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ++cnt;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        acc[c].x += v[u][c].x; acc[c].y += v[u][c].y; acc[c].z += v[u][c].z; acc[c].w += v[u][c].w;
      }
    }
  }
  const bool cont_out = last < g.n_rows && level_key(g, last) == cur;
  flush(is_first && cont_in ? 0 : (cont_out ? 1 : -1));
}

// every cell: empty -> zero row; crossing a chunk border -> sum of its chunk partials in chunk order.
// Lane j of a warp inspects cell base + j (two coalesced loads); only the flagged cells -- a small minority on
// fine levels -- are then worked on, RPI at a time by the lane groups, with kFixUnroll partials in flight.
constexpr int kFixUnroll = 4;
constexpr int kLongList = 16;  // partials; longer lists are reduced by the whole warp

template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_reduce_fix_kernel(SegGeom g, int mean, int cpw, const float* __restrict__ scratch, float* __restrict__ plane) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t base = ((int64_t)blockIdx.x * kSegWarps + warp) * cpw;  // cpw cells per warp (coarse levels: fewer)
  if (base >= g.n_seg) return;
  int my_beg = 0, my_end = 0;
  bool need = false;
  if (lane < cpw && base + lane < g.n_seg) {
    my_beg = __ldg(g.cell_start + ((base + lane) << g.shift));
    my_end = __ldg(g.cell_start + ((base + lane + 1) << g.shift));
    need = my_beg < my_end && my_beg / kChunk != (my_end - 1) / kChunk;
    if (my_beg == my_end) {  // empty cell (the common case on fine levels): its own lane writes the zero row
      float* z = plane + plane_row(g, base + lane) * C;
      for (int c4 = 0; c4 < C / 4; ++c4) st4(z + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  // a cell with a long list of partials (a facade: thousands of rows) is summed by ALL lane groups of the warp,
  // each taking a contiguous part of the list, combined in group order; the others one cell per group
  const bool is_long = need && (my_end - 1) / kChunk - my_beg / kChunk >= kLongList;
  const unsigned flagged = __ballot_sync(0xffffffffu, need && !is_long);
  const unsigned longs = __ballot_sync(0xffffffffu, is_long);
  auto sum_range = [&](float4 (&acc)[CH], int beg, int ca, int cb) {  // partials of chunks [ca, cb] of the cell starting at beg
    for (int ck = ca; ck <= cb; ck += kFixUnroll) {
      float4 o[kFixUnroll][CH];
#pragma unroll
      for (int q = 0; q < kFixUnroll; ++q) {
        const int cq = min(ck + q, cb);
        const float* sp = scratch + ((int64_t)cq * 2 + ((cq * kChunk > beg) ? 0 : 1)) * C + l * 4;
#pragma unroll
        for (int c = 0; c < CH; ++c) o[q][c] = ld4(sp + c * LPR * 4);
      }
#pragma unroll
      for (int q = 0; q < kFixUnroll; ++q)
        if (ck + q <= cb) {
#pragma unroll
          for (int c = 0; c < CH; ++c) { acc[c].x += o[q][c].x; acc[c].y += o[q][c].y; acc[c].z += o[q][c].z; acc[c].w += o[q][c].w; }
        }
    }
  };
  auto finish = [&](float4 (&acc)[CH], int beg, int end, int64_t seg) {
    const float inv = mean ? __fdiv_rn(1.0f, (float)(end - beg)) : 1.0f;
    float* dst = plane + plane_row(g, seg) * C + l * 4;
#pragma unroll
    for (int c = 0; c < CH; ++c) st4(dst + c * LPR * 4, make_float4(acc[c].x * inv, acc[c].y * inv, acc[c].z * inv, acc[c].w * inv));
  };
  const int cnt = __popc(flagged);
  for (int t0 = 0; t0 < cnt; t0 += RPI) {
    const int k = t0 + sub;
    const bool act = k < cnt;
    const int src = act ? (int)__fns(flagged, 0, k + 1) : 0;
    const int beg = __shfl_sync(0xffffffffu, my_beg, src), end = __shfl_sync(0xffffffffu, my_end, src);
    if (!act) continue;
    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    sum_range(acc, beg, beg / kChunk, (end - 1) / kChunk);
    finish(acc, beg, end, base + src);
  }
  for (unsigned rest = longs; rest; rest &= rest - 1) {
    const int src = __ffs(rest) - 1;
    const int beg = __shfl_sync(0xffffffffu, my_beg, src), end = __shfl_sync(0xffffffffu, my_end, src);
    const int c0 = beg / kChunk, c1 = (end - 1) / kChunk;
    const int per = (c1 - c0 + RPI) / RPI;  // chunks per lane group
    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    sum_range(acc, beg, c0 + sub * per, min(c0 + (sub + 1) * per - 1, c1));
#pragma unroll
    for (int off = LPR; off < kWarp; off <<= 1)
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 o = shfl_xor4(acc[c], off);
        acc[c].x += o.x; acc[c].y += o.y; acc[c].z += o.z; acc[c].w += o.w;
      }
    if (sub == 0) finish(acc, beg, end, base + src);
  }
}

// ---- S1 forward: segment max + argmax ---------------------------------------------------------------
// scratch: values [n_chunks][2][C] floats, then positions [n_chunks][2][C] int32 (sorted positions)
template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_max_rows_kernel(const float* __restrict__ rows, SegGeom g, float* __restrict__ plane, int32_t* __restrict__ arg,
                    float* __restrict__ s_val, int32_t* __restrict__ s_pos) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  constexpr int U = CH >= 4 ? 2 : 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t chunk = ((int64_t)blockIdx.x * kSegWarps + warp) * RPI + sub;
  const int64_t first = chunk * kChunk;
  if (first >= g.n_rows) return;
  const int64_t last = min(first + (int64_t)kChunk, g.n_rows);

  int cur = level_key(g, first);
  const bool cont_in = first > 0 && level_key(g, first - 1) == cur;
  bool is_first = true;
  float best[CH][4];
  int bpos[CH][4];
  auto reset = [&]() {
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) { best[c][k] = -FLT_MAX; bpos[c][k] = INT32_MAX; }
  };
  reset();
  auto flush = [&](int slot) {
    if (slot < 0) {
      const int64_t o = plane_row(g, cur) * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        float4 v;
        int4 a;
        float* vp = &v.x;
        int* ap = &a.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool none = bpos[c][k] == INT32_MAX;
          vp[k] = none ? 0.0f : best[c][k];
          ap[k] = none ? -1 : (g.perm ? __ldg(g.perm + bpos[c][k]) : bpos[c][k]);
        }
        if (plane) st4(plane + o + c * LPR * 4, v);
        *reinterpret_cast<int4*>(arg + o + c * LPR * 4) = a;
      }
    } else {
      const int64_t o = (chunk * 2 + slot) * C + l * 4;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        st4(s_val + o + c * LPR * 4, make_float4(best[c][0], best[c][1], best[c][2], best[c][3]));
        *reinterpret_cast<int4*>(s_pos + o + c * LPR * 4) = make_int4(bpos[c][0], bpos[c][1], bpos[c][2], bpos[c][3]);
      }
    }
  };

  for (int64_t i = first; i < last; i += U) {
    int k[U];
    float4 v[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = min(i + u, last - 1);
      k[u] = level_key(g, p);
      const int64_t row = g.perm ? (int64_t)__ldg(g.perm + p) : p;
#pragma unroll
      for (int c = 0; c < CH; ++c) v[u][c] = ld4(rows + row * C + (c * LPR + l) * 4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u >= last) break;
      if (k[u] != cur) {
        flush(is_first && cont_in ? 0 : -1);
        is_first = false;
        cur = k[u];
        reset();
      }
      const int pos = (int)(i + u);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float e[4] = {v[u][c].x, v[u][c].y, v[u][c].z, v[u][c].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // strict >: the first row wins ties (rows of one fine cell are in point order); inside a
          // coarser segment equal values are resolved by the original point index
          bool take = e[q] > best[c][q];
          if (g.tie && e[q] == best[c][q] && bpos[c][q] != INT32_MAX) take = __ldg(g.tie + pos) < __ldg(g.tie + bpos[c][q]);
          if (take) { best[c][q] = e[q]; bpos[c][q] = pos; }
        }
      }
    }
  }
  const bool cont_out = last < g.n_rows && level_key(g, last) == cur;
  flush(is_first && cont_in ? 0 : (cont_out ? 1 : -1));
}

template <class RS>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_max_fix_kernel(SegGeom g, int cpw, const float* __restrict__ s_val, const int32_t* __restrict__ s_pos,
                   float* __restrict__ plane, int32_t* __restrict__ arg) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t base = ((int64_t)blockIdx.x * kSegWarps + warp) * cpw;
  if (base >= g.n_seg) return;
  int my_beg = 0, my_end = 0;
  bool need = false;
  if (lane < cpw && base + lane < g.n_seg) {
    my_beg = __ldg(g.cell_start + ((base + lane) << g.shift));
    my_end = __ldg(g.cell_start + ((base + lane + 1) << g.shift));
    need = my_beg < my_end && my_beg / kChunk != (my_end - 1) / kChunk;
    if (my_beg == my_end) {  // empty cell: value 0, arg -1, written by its own lane
      const int64_t z = plane_row(g, base + lane) * C;
      for (int c4 = 0; c4 < C / 4; ++c4) {
        if (plane) st4(plane + z + c4 * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        *reinterpret_cast<int4*>(arg + z + c4 * 4) = make_int4(-1, -1, -1, -1);
      }
    }
  }
  const bool is_long = need && (my_end - 1) / kChunk - my_beg / kChunk >= kLongList;
  const unsigned flagged = __ballot_sync(0xffffffffu, need && !is_long);
  const unsigned longs = __ballot_sync(0xffffffffu, is_long);
  // does candidate (ov, op) beat (bv, bp)?  larger value; equal values -> earlier original point
  auto beats = [&](float ov, int op, float bv, int bp) -> bool {
    if (ov > bv) return true;
    if (ov == bv && op != INT32_MAX && bp != INT32_MAX) return g.tie ? (__ldg(g.tie + op) < __ldg(g.tie + bp)) : (op < bp);
    return false;
  };
  auto max_range = [&](float (&best)[4], int (&bpos)[4], int c, int beg, int ca, int cb) {
    for (int ck = ca; ck <= cb; ck += kFixUnroll) {
      float4 ov4[kFixUnroll];
      int4 op4[kFixUnroll];
#pragma unroll
      for (int q = 0; q < kFixUnroll; ++q) {
        const int cq = min(ck + q, cb);
        const int64_t so = ((int64_t)cq * 2 + ((cq * kChunk > beg) ? 0 : 1)) * C + (c * LPR + l) * 4;
        ov4[q] = ld4(s_val + so);
        op4[q] = __ldg(reinterpret_cast<const int4*>(s_pos + so));
      }
#pragma unroll
      for (int q = 0; q < kFixUnroll; ++q) {
        if (ck + q > cb) break;
        const float ov[4] = {ov4[q].x, ov4[q].y, ov4[q].z, ov4[q].w};
        const int op[4] = {op4[q].x, op4[q].y, op4[q].z, op4[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (beats(ov[e], op[e], best[e], bpos[e])) { best[e] = ov[e]; bpos[e] = op[e]; }
      }
    }
  };
  auto emit = [&](const float (&best)[4], const int (&bpos)[4], int c, int64_t seg) {
    const int64_t o = plane_row(g, seg) * C + (c * LPR + l) * 4;
    float4 v;
    int4 a;
    float* vp = &v.x;
    int* ap = &a.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool none = bpos[e] == INT32_MAX;
      vp[e] = none ? 0.0f : best[e];
      ap[e] = none ? -1 : (g.perm ? __ldg(g.perm + bpos[e]) : bpos[e]);
    }
    if (plane) st4(plane + o, v);
    *reinterpret_cast<int4*>(arg + o) = a;
  };
  const int cnt = __popc(flagged);
  for (int t0 = 0; t0 < cnt; t0 += RPI) {
    const int k = t0 + sub;
    const bool act = k < cnt;
    const int src = act ? (int)__fns(flagged, 0, k + 1) : 0;
    const int beg = __shfl_sync(0xffffffffu, my_beg, src), end = __shfl_sync(0xffffffffu, my_end, src);
    if (!act) continue;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float best[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
      int bpos[4] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX};
      max_range(best, bpos, c, beg, beg / kChunk, (end - 1) / kChunk);
      emit(best, bpos, c, base + src);
    }
  }
  for (unsigned rest = longs; rest; rest &= rest - 1) {
    const int src = __ffs(rest) - 1;
    const int beg = __shfl_sync(0xffffffffu, my_beg, src), end = __shfl_sync(0xffffffffu, my_end, src);
    const int c0 = beg / kChunk, c1 = (end - 1) / kChunk;
    const int per = (c1 - c0 + RPI) / RPI;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float best[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
      int bpos[4] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX};
      max_range(best, bpos, c, beg, c0 + sub * per, min(c0 + (sub + 1) * per - 1, c1));
      // the (value, original point) order is total, so the combination order does not matter
#pragma unroll
      for (int off = LPR; off < kWarp; off <<= 1)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ov = __shfl_xor_sync(0xffffffffu, best[e], off);
          const int op = __shfl_xor_sync(0xffffffffu, bpos[e], off);
          if (beats(ov, op, best[e], bpos[e])) { best[e] = ov; bpos[e] = op; }
        }
      if (sub == 0) emit(best, bpos, c, base + src);
    }
  }
}

// ---- row-parallel maps: every point looks up the plane row of its cell ------------------------------
// A warp owns 32 consecutive sorted positions per pass: lane j resolves (plane row, 1/count, row) of position
// j once, the RPI sub-groups fetch them by shuffle.
//   MODE 0: rows[row] = plane[cell] (* 1/count when mean)       -- S2 backward, gather-back of pool_local
//   MODE 1: rows[row, c] = (arg[cell, c] == row) ? tot[cell, c] (+ extra[cell, c]) : 0   -- S1 backward
//   MODE 2: rows[row] = plane[cell] (* 1/count) + extra[row], *slot = max |rows|  -- S2 backward fused with the
//           accumulation of the second gradient branch of the same tensor (and the operand maximum of the GEMMs
//           that consume the sum)
template <class RS, int MODE>
__global__ void __launch_bounds__(kSegWarps * kWarp)
seg_rowmap_kernel(const float* __restrict__ plane, const float* __restrict__ extra, const int32_t* __restrict__ arg,
                  SegGeom g, int mean, float* __restrict__ rows, uint32_t* __restrict__ slot = nullptr) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t first = ((int64_t)blockIdx.x * kSegWarps + (threadIdx.x >> 5)) * 32;
  if (first >= g.n_rows) return;
  const int npts = (int)min((int64_t)32, g.n_rows - first);
  int64_t my_prow = 0;
  int my_row = 0;
  float my_inv = 1.0f;
  if (lane < npts) {
    const int64_t i = first + lane;
    const int64_t seg = level_key(g, i);
    my_prow = plane_row(g, seg);
    my_row = g.perm ? __ldg(g.perm + i) : (int)i;
    if (mean) {
      const int n = __ldg(g.cell_start + ((seg + 1) << g.shift)) - __ldg(g.cell_start + (seg << g.shift));
      my_inv = __fdiv_rn(1.0f, (float)n);
    }
  }
  float vmax = 0.f;
  for (int t0 = 0; t0 < npts; t0 += RPI) {
    const int j = t0 + sub;
    const bool act = j < npts;
    const int src = act ? j : 0;
    const int64_t prow = __shfl_sync(0xffffffffu, my_prow, src);
    const int row = __shfl_sync(0xffffffffu, my_row, src);
    const float inv = __shfl_sync(0xffffffffu, my_inv, src);
    if (!act) continue;
    const float* p = plane + prow * C + l * 4;
    float* dst = rows + (int64_t)row * C + l * 4;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      float4 v = ld4(p + c * LPR * 4);
      if (MODE == 0) {
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      } else if (MODE == 2) {
        const float4 e = ld4_stream(extra + (int64_t)row * C + (c * LPR + l) * 4);
        v.x = v.x * inv + e.x; v.y = v.y * inv + e.y; v.z = v.z * inv + e.z; v.w = v.w * inv + e.w;
        vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      } else {
        if (extra) {
          const float4 e = ld4(extra + prow * C + (c * LPR + l) * 4);
          v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        const int4 a = __ldg(reinterpret_cast<const int4*>(arg + prow * C + (c * LPR + l) * 4));
        v.x = a.x == row ? v.x : 0.f;
        v.y = a.y == row ? v.y : 0.f;
        v.z = a.z == row ? v.z : 0.f;
        v.w = a.w == row ? v.w : 0.f;
      }
      st4_stream(dst + c * LPR * 4, v);
    }
  }
  if (MODE == 2 && slot) {  // bit pattern order == float order for non-negative values; integer atomicMax is exact
    uint32_t b = __float_as_uint(vmax);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane == 0 && b) atomicMax(slot, b);
  }
}

static int check_geom(const int32_t* row_keys, const int32_t* cell_start, int64_t n_rows, int64_t n_seg, int shift, int C,
                      int morton, int reso) {
  if (!cell_start || n_rows < 0 || n_seg < 0 || shift < 0 || shift > 30 || (shift & 1) || C <= 0 || reso <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if (n_rows > 0 && !row_keys) return T2H_ERR_INVALID_ARGUMENT;
  if (n_rows > INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;  // positions are int32 (cell_start is)
  if (shift && !morton) return T2H_ERR_INVALID_ARGUMENT;  // only Morton keys nest across levels
  if (morton && ((reso & (reso - 1)) || n_seg % ((int64_t)reso * reso))) return T2H_ERR_INVALID_ARGUMENT;
  return T2H_OK;
}

static inline int log2_cells_of(int reso) {
  int l = 0;
  while ((1 << l) < reso) ++l;
  return 2 * l;
}

static inline int64_t n_chunks_of(int64_t n_rows) { return (n_rows + kChunk - 1) / kChunk; }
static inline unsigned blocks_for(int64_t items, int per_warp) {
  const int64_t warps = (items + per_warp - 1) / per_warp;
  return (unsigned)((warps + kSegWarps - 1) / kSegWarps);
}
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// cells inspected per warp of a fix-up launch: 32 on fine levels (most cells need nothing), fewer on coarse
// levels (every cell sums several partials) so that the launch still fills the machine
static inline int fix_cells_per_warp(int64_t n_seg, int rpi) {
  int cpw = 32;
  while (cpw > rpi && n_seg / cpw < 8192) cpw >>= 1;
  return cpw;
}

}  // namespace t2h

using namespace t2h;

extern "C" size_t t2h_seg_workspace_bytes(int64_t n_rows, int64_t n_seg, int C) {
  // chunk-border partials: values + positions (max), or sums (mean); plus one (n_seg, C) plane of segment
  // sums for the backward of the max
  const size_t slots = (size_t)n_chunks_of(n_rows > 0 ? n_rows : 0) * 2 * (size_t)C;
  return align256(slots * sizeof(float)) + align256(slots * sizeof(int32_t)) + align256((size_t)(n_seg > 0 ? n_seg : 0) * C * sizeof(float)) + 256;
}

extern "C" int t2h_seg_max_fwd(const float* rows, int64_t n_rows, const int32_t* perm, const int32_t* tie_rank,
                               const int32_t* row_keys, const int32_t* cell_start, int64_t n_seg, int shift, int C,
                               int morton, int reso, void* workspace, size_t workspace_bytes, float* pooled, float* plane,
                               int32_t* arg, t2h_stream_t stream) {
  int st = check_geom(row_keys, cell_start, n_rows, n_seg, shift, C, morton, reso);
  if (st) return st;
  if ((n_rows > 0 && !rows) || !arg || (pooled && !plane)) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  if (!workspace || workspace_bytes < t2h_seg_workspace_bytes(n_rows, n_seg, C)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  SegGeom g{perm, tie_rank, row_keys, cell_start, n_rows, n_seg, shift, morton, reso, log2_cells_of(reso)};
  const size_t slots = (size_t)n_chunks_of(n_rows) * 2 * (size_t)C;
  float* s_val = (float*)workspace;
  int32_t* s_pos = (int32_t*)((char*)workspace + align256(slots * sizeof(float)));
  cudaStream_t s = (cudaStream_t)stream;
  T2H_DISPATCH_ROWSHAPE(C, {
    if (n_rows > 0)
      seg_max_rows_kernel<RS><<<blocks_for(n_chunks_of(n_rows), RS::RPI), kSegWarps * kWarp, 0, s>>>(rows, g, plane, arg, s_val, s_pos);
    const int cpw = fix_cells_per_warp(n_seg, RS::RPI);
    seg_max_fix_kernel<RS><<<blocks_for(n_seg, cpw), kSegWarps * kWarp, 0, s>>>(g, cpw, s_val, s_pos, plane, arg);
    if (pooled && n_rows > 0)
      seg_rowmap_kernel<RS, 0><<<blocks_for(n_rows, 32), kSegWarps * kWarp, 0, s>>>(plane, nullptr, nullptr, g, 0, pooled);
  });
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_max_bwd(const float* grad_pooled, const float* grad_plane, int64_t n_rows, const int32_t* perm,
                               const int32_t* row_keys, const int32_t* cell_start, int64_t n_seg, int shift, int C,
                               int morton, int reso, const int32_t* arg, void* workspace, size_t workspace_bytes,
                               float* grad_rows, t2h_stream_t stream) {
  int st = check_geom(row_keys, cell_start, n_rows, n_seg, shift, C, morton, reso);
  if (st) return st;
  if (!arg || (n_rows > 0 && !grad_rows) || (!grad_pooled && !grad_plane)) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0 || n_rows == 0) return T2H_OK;
  SegGeom g{perm, nullptr, row_keys, cell_start, n_rows, n_seg, shift, morton, reso, log2_cells_of(reso)};
  cudaStream_t s = (cudaStream_t)stream;
  const float* tot = grad_plane;
  const float* extra = nullptr;
  if (grad_pooled) {
    if (!workspace || workspace_bytes < t2h_seg_workspace_bytes(n_rows, n_seg, C)) return T2H_ERR_WORKSPACE_TOO_SMALL;
    const size_t slots = (size_t)n_chunks_of(n_rows) * 2 * (size_t)C;
    float* scratch = (float*)workspace;
    float* sums = (float*)((char*)workspace + align256(slots * sizeof(float)) + align256(slots * sizeof(int32_t)));
    T2H_DISPATCH_ROWSHAPE(C, {
      seg_reduce_rows_kernel<RS><<<blocks_for(n_chunks_of(n_rows), RS::RPI), kSegWarps * kWarp, 0, s>>>(grad_pooled, g, 0, sums, scratch);
      const int cpw = fix_cells_per_warp(n_seg, RS::RPI);
      seg_reduce_fix_kernel<RS><<<blocks_for(n_seg, cpw), kSegWarps * kWarp, 0, s>>>(g, 0, cpw, scratch, sums);
    });
    tot = sums;
    extra = grad_plane;
  }
  T2H_DISPATCH_ROWSHAPE(C, (seg_rowmap_kernel<RS, 1><<<blocks_for(n_rows, 32), kSegWarps * kWarp, 0, s>>>(tot, extra, arg, g, 0, grad_rows)));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_reduce_fwd(const float* rows, int64_t n_rows, const int32_t* perm, const int32_t* row_keys,
                                  const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                                  int mean, void* workspace, size_t workspace_bytes, float* plane, t2h_stream_t stream) {
  int st = check_geom(row_keys, cell_start, n_rows, n_seg, shift, C, morton, reso);
  if (st) return st;
  if ((n_rows > 0 && !rows) || !plane) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  if (!workspace || workspace_bytes < t2h_seg_workspace_bytes(n_rows, n_seg, C)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  SegGeom g{perm, nullptr, row_keys, cell_start, n_rows, n_seg, shift, morton, reso, log2_cells_of(reso)};
  float* scratch = (float*)workspace;
  cudaStream_t s = (cudaStream_t)stream;
  T2H_DISPATCH_ROWSHAPE(C, {
    if (n_rows > 0)
      seg_reduce_rows_kernel<RS><<<blocks_for(n_chunks_of(n_rows), RS::RPI), kSegWarps * kWarp, 0, s>>>(rows, g, mean, plane, scratch);
    const int cpw = fix_cells_per_warp(n_seg, RS::RPI);
    seg_reduce_fix_kernel<RS><<<blocks_for(n_seg, cpw), kSegWarps * kWarp, 0, s>>>(g, mean, cpw, scratch, plane);
  });
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_broadcast(const float* plane, int64_t n_rows, const int32_t* perm, const int32_t* row_keys,
                                 const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                                 int mean, float* rows, t2h_stream_t stream) {
  int st = check_geom(row_keys, cell_start, n_rows, n_seg, shift, C, morton, reso);
  if (st) return st;
  if ((n_rows > 0 && !rows) || !plane) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0 || n_rows == 0) return T2H_OK;
  SegGeom g{perm, nullptr, row_keys, cell_start, n_rows, n_seg, shift, morton, reso, log2_cells_of(reso)};
  T2H_DISPATCH_ROWSHAPE(C, (seg_rowmap_kernel<RS, 0><<<blocks_for(n_rows, 32), kSegWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               plane, nullptr, nullptr, g, mean, rows)));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_seg_broadcast_add(const float* plane, const float* add_rows, int64_t n_rows, const int32_t* perm,
                                     const int32_t* row_keys, const int32_t* cell_start, int64_t n_seg, int shift, int C,
                                     int morton, int reso, int mean, float* rows, uint32_t* absmax_slot,
                                     t2h_stream_t stream) {
  int st = check_geom(row_keys, cell_start, n_rows, n_seg, shift, C, morton, reso);
  if (st) return st;
  if ((n_rows > 0 && (!rows || !add_rows)) || !plane) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0 || n_rows == 0) return T2H_OK;
  SegGeom g{perm, nullptr, row_keys, cell_start, n_rows, n_seg, shift, morton, reso, log2_cells_of(reso)};
  if (absmax_slot && cudaMemsetAsync(absmax_slot, 0, sizeof(uint32_t), (cudaStream_t)stream) != cudaSuccess) return T2H_ERR_CUDA;
  T2H_DISPATCH_ROWSHAPE(C, (seg_rowmap_kernel<RS, 2><<<blocks_for(n_rows, 32), kSegWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               plane, add_rows, nullptr, g, mean, rows, absmax_slot)));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

/* ---- aliases named in SURVEY.md §8(b): the mean-scatter pair ---------------------------------------- */
extern "C" int t2h_seg_mean_fwd(const float* rows, int64_t n_rows, const int32_t* perm, const int32_t* row_keys,
                                const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                                void* workspace, size_t workspace_bytes, float* plane, t2h_stream_t stream) {
  return t2h_seg_reduce_fwd(rows, n_rows, perm, row_keys, cell_start, n_seg, shift, C, morton, reso, 1, workspace,
                            workspace_bytes, plane, stream);
}
extern "C" int t2h_seg_mean_bwd(const float* grad_plane, int64_t n_rows, const int32_t* perm, const int32_t* row_keys,
                                const int32_t* cell_start, int64_t n_seg, int shift, int C, int morton, int reso,
                                float* grad_rows, t2h_stream_t stream) {
  return t2h_seg_broadcast(grad_plane, n_rows, perm, row_keys, cell_start, n_seg, shift, C, morton, reso, 1, grad_rows, stream);
}
