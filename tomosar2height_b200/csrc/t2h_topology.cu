// Cell index (a1) and the point->cell topology: keys, stable sort by cell, cell_start table.
//
// Replaces utils/coordinate.py:12-28 (coordinate2index) and its 10-12 re-evaluations per
// forward (pointnet.py:70, alto.py:80,190): the points are sorted ONCE by the Morton code of
// their finest cell, after which the segments of every power-of-two plane resolution are
// contiguous key ranges (see include/t2h.h).
#include "t2h_common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace t2h {

__global__ void cell_index_kernel(const float* __restrict__ xy, int64_t n, int64_t stride, int reso,
                                  int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = (float)reso;
  // (x * reso).long(): fp32 multiply (round-to-nearest), then truncation toward zero
  long long ix = (long long)__fmul_rn(xy[i * stride], r);
  long long iy = (long long)__fmul_rn(xy[i * stride + 1], r);
  out[i] = ix + (long long)reso * iy;
}

__global__ void xy_keys_kernel(const float* __restrict__ xyz, int64_t n, int64_t stride,
                               int64_t n_per_batch, int reso, int morton, int32_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = (float)reso;
  int ix = __float2int_rz(__fmul_rn(xyz[i * stride], r));
  int iy = __float2int_rz(__fmul_rn(xyz[i * stride + 1], r));
  ix = min(max(ix, 0), reso - 1);  // the reference relies on the dataset crop (dataset.py:278)
  iy = min(max(iy, 0), reso - 1);
  int64_t b = i / n_per_batch;
  keys[i] = (int32_t)(b * (int64_t)reso * reso + cell_code((uint32_t)ix, (uint32_t)iy, reso, morton));
}

// ragged batches: tile b owns the points [offsets[b], offsets[b+1]) of the flat cloud
__global__ void xy_keys_ragged_kernel(const float* __restrict__ xyz, int64_t n, int64_t stride,
                                      const int64_t* __restrict__ offsets, int n_tiles, int reso, int morton,
                                      int32_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_tiles;  // largest b with offsets[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid;
  }
  const float r = (float)reso;
  int ix = __float2int_rz(__fmul_rn(xyz[i * stride], r));
  int iy = __float2int_rz(__fmul_rn(xyz[i * stride + 1], r));
  ix = min(max(ix, 0), reso - 1);
  iy = min(max(iy, 0), reso - 1);
  keys[i] = (int32_t)((int64_t)lo * reso * reso + cell_code((uint32_t)ix, (uint32_t)iy, reso, morton));
}

__global__ void index_keys_kernel(const int64_t* __restrict__ index, int64_t n, int64_t n_per_batch,
                                  int64_t dim_size, int32_t* __restrict__ keys, int32_t* flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = index[i];
  if (v < 0 || v >= dim_size) { *flag = 1; v = v < 0 ? 0 : dim_size - 1; }
  keys[i] = (int32_t)((i / n_per_batch) * dim_size + v);
}

__global__ void iota_kernel(int32_t* __restrict__ p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

// cell_start[k] = first sorted position whose key is >= k.  Thread i owns the keys in
// (key[i-1], key[i]]; thread n owns (key[n-1], n_keys].
__global__ void cell_start_kernel(const int32_t* __restrict__ keys_sorted, int64_t n, int64_t n_keys,
                                  int32_t* __restrict__ cell_start) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int64_t lo = (i == 0) ? -1 : (int64_t)keys_sorted[i - 1];
  int64_t hi = (i == n) ? n_keys : (int64_t)keys_sorted[i];
  for (int64_t k = lo + 1; k <= hi; ++k) cell_start[k] = (int32_t)i;
}

template <int W>
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm,
                                   int64_t n_rows, int width, float* __restrict__ dst, int scatter) {
  // one thread per (row, W-float group)
  int groups = width / W;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * groups) return;
  int64_t i = t / groups;
  int g = (int)(t % groups);
  int64_t s = scatter ? i : (int64_t)perm[i];
  int64_t d = scatter ? (int64_t)perm[i] : i;
  if (W == 4) {
    st4(dst + d * width + g * 4, ld4(src + s * width + g * 4));
  } else {
    dst[d * width + g] = src[s * width + g];
  }
}

static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

static size_t cub_temp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  return bytes;
}

}  // namespace t2h

using namespace t2h;

extern "C" int t2h_cell_index(const float* xy, int64_t n_points, int64_t point_stride, int reso,
                              int64_t* out_index, t2h_stream_t stream) {
  if (!xy || !out_index || n_points < 0 || point_stride < 2 || reso <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (n_points == 0) return T2H_OK;
  cell_index_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xy, n_points, point_stride, reso, out_index);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_xy_keys(const float* xyz, int64_t n_points, int64_t point_stride, int64_t n_per_batch,
                           int reso, int morton, int32_t* keys, t2h_stream_t stream) {
  if (!xyz || !keys || n_points < 0 || point_stride < 2 || n_per_batch <= 0 || reso <= 0 || reso > 32768)
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;  // Morton keys need a power of two
  int64_t n_batch = (n_points + n_per_batch - 1) / n_per_batch;
  if (n_batch * (int64_t)reso * reso > (int64_t)INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (n_points == 0) return T2H_OK;
  xy_keys_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n_points, point_stride, n_per_batch, reso, morton, keys);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_xy_keys_ragged(const float* xyz, int64_t n_points, int64_t point_stride, const int64_t* offsets,
                                  int n_tiles, int reso, int morton, int32_t* keys, t2h_stream_t stream) {
  if (!xyz || !keys || !offsets || n_points < 0 || point_stride < 2 || n_tiles <= 0 || reso <= 0 || reso > 32768)
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;
  if ((int64_t)n_tiles * reso * reso > (int64_t)INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (n_points == 0) return T2H_OK;
  xy_keys_ragged_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n_points, point_stride, offsets,
                                                                                    n_tiles, reso, morton, keys);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_index_keys(const int64_t* index, int64_t n_points, int64_t n_per_batch, int64_t dim_size,
                              int32_t* keys, int32_t* flag, t2h_stream_t stream) {
  if (!index || !keys || !flag || n_points < 0 || n_per_batch <= 0 || dim_size <= 0) return T2H_ERR_INVALID_ARGUMENT;
  int64_t n_batch = (n_points + n_per_batch - 1) / n_per_batch;
  if (n_batch * dim_size > (int64_t)INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (n_points == 0) return T2H_OK;
  index_keys_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(index, n_points, n_per_batch, dim_size, keys, flag);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" size_t t2h_sort_workspace_bytes(int64_t n_points) {
  if (n_points <= 0) return 256;
  size_t iota = ((size_t)n_points * sizeof(int32_t) + 255) & ~(size_t)255;
  return iota + cub_temp_bytes(n_points) + 256;
}

extern "C" int t2h_sort_by_cell(const int32_t* keys, int64_t n_points, int64_t n_keys, void* workspace,
                                size_t workspace_bytes, int32_t* keys_sorted, int32_t* perm,
                                int32_t* cell_start, t2h_stream_t stream) {
  if (!keys || !keys_sorted || !perm || !cell_start || !workspace || n_points < 0 || n_keys <= 0 ||
      n_points > (int64_t)INT32_MAX - 1 || n_keys > (int64_t)INT32_MAX - 1)
    return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_sort_workspace_bytes(n_points)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_points > 0) {
    size_t iota_bytes = ((size_t)n_points * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t* iota = (int32_t*)workspace;
    void* temp = (char*)workspace + iota_bytes;
    size_t temp_bytes = workspace_bytes - iota_bytes;
    iota_kernel<<<blocks_for(n_points, 256), 256, 0, s>>>(iota, n_points);
    T2H_CHECK_LAUNCH();
    int end_bit = 1;
    while (end_bit < 31 && ((int64_t)1 << end_bit) < n_keys) ++end_bit;
    // LSD radix sort: stable, so equal keys keep their input (point index) order
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_sorted, (const int32_t*)iota, perm,
                                                    (int)n_points, 0, end_bit, s);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return T2H_ERR_CUDA; }
  }
  cell_start_kernel<<<blocks_for(n_points + 1, 256), 256, 0, s>>>(keys_sorted, n_points, n_keys, cell_start);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

static int rows_copy(const float* src, const int32_t* perm, int64_t n_rows, int width, float* dst, int scatter,
                     t2h_stream_t stream) {
  if (!src || !perm || !dst || n_rows < 0 || width <= 0) return T2H_ERR_INVALID_ARGUMENT;
  if (n_rows == 0) return T2H_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (width % 4 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) {
    int64_t t = n_rows * (width / 4);
    gather_rows_kernel<4><<<blocks_for(t, 256), 256, 0, s>>>(src, perm, n_rows, width, dst, scatter);
  } else {
    int64_t t = n_rows * width;
    gather_rows_kernel<1><<<blocks_for(t, 256), 256, 0, s>>>(src, perm, n_rows, width, dst, scatter);
  }
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_gather_rows(const float* src, const int32_t* perm, int64_t n_rows, int width, float* dst,
                               t2h_stream_t stream) {
  return rows_copy(src, perm, n_rows, width, dst, 0, stream);
}

extern "C" int t2h_scatter_rows(const float* src, const int32_t* perm, int64_t n_rows, int width, float* dst,
                                t2h_stream_t stream) {
  return rows_copy(src, perm, n_rows, width, dst, 1, stream);
}

extern "C" int t2h_abi_version(void) { return 1; }

extern "C" const char* t2h_status_string(int status) {
  switch (status) {
    case T2H_OK: return "ok";
    case T2H_ERR_INVALID_ARGUMENT: return "invalid argument";
    case T2H_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
    case T2H_ERR_CUDA: return "CUDA error";
    case T2H_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    default: return "unknown status";
  }
}
