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

template <class RS>
__global__ void __launch_bounds__(kSampleWarps * kWarp)
sample_fwd_kernel(const float* __restrict__ plane, int reso, const float* __restrict__ xyz, int64_t stride,
                  const int32_t* __restrict__ perm, int64_t n, int64_t n_per_batch, float* __restrict__ out) {
  constexpr int LPR = RS::LPR, CH = RS::CH, RPI = RS::RPI, C = RS::C;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t warp = (int64_t)blockIdx.x * kSampleWarps + (threadIdx.x >> 5);
  const int64_t first = warp * kPointsPerWarp;
  const int64_t last = min(first + kPointsPerWarp, n);
  for (int64_t i = first + sub; i < last; i += RPI) {
    const int64_t row = perm ? (int64_t)perm[i] : i;
    const int64_t b = row / n_per_batch;
    const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + i * stride));
    const Taps t = make_taps(pxy.x, pxy.y, reso);
    // taps beyond the last row/column carry zero weight (ix == r-1 exactly); clamp the address
    const int x1 = min(t.x0 + 1, reso - 1), y1 = min(t.y0 + 1, reso - 1);
    const float w_nw = __fmul_rn(t.wx0, t.wy0);
    const float w_ne = (t.x0 + 1 < reso) ? __fmul_rn(t.wx1, t.wy0) : 0.f;
    const float w_sw = (t.y0 + 1 < reso) ? __fmul_rn(t.wx0, t.wy1) : 0.f;
    const float w_se = (t.x0 + 1 < reso && t.y0 + 1 < reso) ? __fmul_rn(t.wx1, t.wy1) : 0.f;
    const float* base = plane + (b * reso * (int64_t)reso) * C + l * 4;
    const float* p_nw = base + ((int64_t)t.y0 * reso + t.x0) * C;
    const float* p_ne = base + ((int64_t)t.y0 * reso + x1) * C;
    const float* p_sw = base + ((int64_t)y1 * reso + t.x0) * C;
    const float* p_se = base + ((int64_t)y1 * reso + x1) * C;
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
        for (int i = beg + sub; i < end; i += RPI) {
          const float2 pxy = __ldg(reinterpret_cast<const float2*>(xyz + (int64_t)i * stride));
          const Taps t = make_taps(pxy.x, pxy.y, reso);
          const float wx = (t.x0 == cx) ? t.wx0 : ((t.x0 + 1 == cx) ? t.wx1 : 0.f);
          const float wy = (t.y0 == cy) ? t.wy0 : ((t.y0 + 1 == cy) ? t.wy1 : 0.f);
          const float w = __fmul_rn(wx, wy);
          if (w != 0.f) {
            const int64_t row = perm ? (int64_t)perm[i] : (int64_t)i;
            const float* src = grad_rows + row * C + l * 4;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const float4 gq = ld4(src + c * LPR * 4);
              acc[c].x += w * gq.x; acc[c].y += w * gq.y; acc[c].z += w * gq.z; acc[c].w += w * gq.w;
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
                                       int64_t point_stride, const int32_t* perm, int64_t n_points,
                                       int64_t n_per_batch, float* out_rows, t2h_stream_t stream) {
  if (!plane || !xyz_sorted || !out_rows || reso <= 0 || point_stride < 2 || (point_stride & 1) || n_points < 0 ||
      n_per_batch <= 0)
    return T2H_ERR_INVALID_ARGUMENT;
  if (n_points == 0) return T2H_OK;
  const int64_t warps = (n_points + kPointsPerWarp - 1) / kPointsPerWarp;
  const unsigned blocks = (unsigned)((warps + kSampleWarps - 1) / kSampleWarps);
  T2H_DISPATCH_ROWSHAPE(C, sample_fwd_kernel<RS><<<blocks, kSampleWarps * kWarp, 0, (cudaStream_t)stream>>>(
                               plane, reso, xyz_sorted, point_stride, perm, n_points, n_per_batch, out_rows));
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_bilinear_sample_bwd(const float* grad_rows, int64_t n_points, int reso, int C,
                                       const float* xyz_sorted, int64_t point_stride, const int32_t* perm,
                                       const int32_t* cell_start, int64_t n_seg, int shift, int morton,
                                       float* grad_plane, t2h_stream_t stream) {
  if (!grad_rows || !xyz_sorted || !cell_start || !grad_plane || reso <= 0 || point_stride < 2 || (point_stride & 1) ||
      n_points < 0 || n_seg < 0 ||
      shift < 0 || (shift & 1) || (shift && !morton) || n_seg % ((int64_t)reso * reso))
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;
  if (n_seg == 0) return T2H_OK;
  // ~9 neighbour cells are scanned per plane cell: split the scan over several warps on coarse levels
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
