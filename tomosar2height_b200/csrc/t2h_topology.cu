// Cell index (a1) and the point->cell topology: keys, stable sort by cell, cell_start table.
//
// Replaces utils/coordinate.py:12-28 (coordinate2index) and its 10-12 re-evaluations per
// forward (pointnet.py:70, alto.py:80,190): the points are sorted ONCE by the Morton code of
// their finest cell, after which the segments of every power-of-two plane resolution are
// contiguous key ranges (see include/t2h.h).
#include "t2h_common.cuh"

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

// cell coordinates of a point, with the reference's arithmetic ((x * reso).long(): coordinate.py:24-26).  The reference
// does not clamp: it relies on the dataset's crop to the open unit square (dataset.py:278), and a point outside it
// makes torch_scatter fail on the index.  Here such a point (incl. NaN) is binned into the nearest border cell so
// that no kernel can run out of bounds, and `*flag` (nullable) is raised so that the caller can fail like the reference.
__device__ __forceinline__ void cell_of(const float* __restrict__ p, int reso, int& ix, int& iy, int32_t* flag) {
  const float r = (float)reso;
  ix = __float2int_rz(__fmul_rn(p[0], r));
  iy = __float2int_rz(__fmul_rn(p[1], r));
  const bool bad = !(p[0] >= 0.f) || !(p[1] >= 0.f) || ix >= reso || iy >= reso;  // !(>=) also catches NaN
  if (bad && flag) *flag = 1;
  ix = min(max(ix, 0), reso - 1);
  iy = min(max(iy, 0), reso - 1);
}

__global__ void xy_keys_kernel(const float* __restrict__ xyz, int64_t n, int64_t stride,
                               int64_t n_per_batch, int reso, int morton, int32_t* __restrict__ keys, int32_t* flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ix, iy;
  cell_of(xyz + i * stride, reso, ix, iy, flag);
  int64_t b = i / n_per_batch;
  keys[i] = (int32_t)(b * (int64_t)reso * reso + cell_code((uint32_t)ix, (uint32_t)iy, reso, morton));
}

// ragged batches: tile b owns the points [offsets[b], offsets[b+1]) of the flat cloud
__global__ void xy_keys_ragged_kernel(const float* __restrict__ xyz, int64_t n, int64_t stride,
                                      const int64_t* __restrict__ offsets, int n_tiles, int reso, int morton,
                                      int32_t* __restrict__ keys, int32_t* flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_tiles;  // largest b with offsets[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid;
  }
  int ix, iy;
  cell_of(xyz + i * stride, reso, ix, iy, flag);
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

// ---- stable LSD radix sort of (key, position) pairs, 8 bits per pass -------------------------------------------
// Keys are cell codes (<= 21 bits for 32 tiles at R = 256), so 2-3 passes.  Per pass:
//   radix_hist_kernel     every CTA histograms the digit of its tile of kSortTile keys (shared-memory integer
//                         atomics -- the counts are exact whatever the order)              -> hist[digit][tile]
//   radix_scan_kernel     CTA d: exclusive scan of hist[d][*] over the tiles, total        -> hist (in place), total[d]
//   radix_scatter_kernel  every CTA re-reads its tile; a warp owns a contiguous piece and ranks its keys round by
//                         round with __match_any_sync (lanes with the same digit, in lane order) on top of a running
//                         per-warp counter, the warps are offset against each other per digit, and the pair goes to
//                         digit base + tile offset + warp offset + rank: equal digits keep their input order (stable),
//                         which is what gives "ties -> smallest point index" further up.
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                        // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kSortWarpKeys = 32 * kSortItems;        // a warp's contiguous piece of the tile

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift, int64_t n_tiles, int32_t* __restrict__ hist) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[((uint32_t)keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
radix_scan_kernel(int32_t* __restrict__ hist, int64_t n_tiles, int32_t* __restrict__ total) {
  __shared__ int part[kSortThreads];
  int32_t* row = hist + (int64_t)blockIdx.x * n_tiles;
  const int64_t per = (n_tiles + kSortThreads - 1) / kSortThreads;
  const int64_t lo = min((int64_t)threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
  int sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += row[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {  // 256 partial sums: serial exclusive scan
    int run = 0;
    for (int t = 0; t < kSortThreads; ++t) { const int v = part[t]; part[t] = run; run += v; }
    total[blockIdx.x] = run;
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int64_t i = lo; i < hi; ++i) { const int v = row[i]; row[i] = run; run += v; }
}

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n, int shift,
                     int64_t n_tiles, const int32_t* __restrict__ hist, const int32_t* __restrict__ total,
                     int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ int wcount[kSortThreads / 32][256];  // per warp: keys of each digit seen so far -> warp offsets
  __shared__ int dbase[256];                      // digit base + this tile's offset inside the digit
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
  {  // exclusive scan of the 256 digit totals (one value per thread)
    __shared__ int tot[256];
    tot[threadIdx.x] = total[threadIdx.x];
    __syncthreads();
    int b = 0;
    for (int d = 0; d < threadIdx.x; ++d) b += tot[d];
    dbase[threadIdx.x] = b + hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x];
  }
  __syncthreads();
  const int64_t first = (int64_t)blockIdx.x * kSortTile + warp * kSortWarpKeys;
  int32_t k[kSortItems], v[kSortItems];
  int rank[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = first + r * 32 + lane;
    const bool ok = i < n;
    k[r] = ok ? keys_in[i] : 0;
    v[r] = ok ? (vals_in ? vals_in[i] : (int32_t)i) : 0;
    // invalid lanes take a private pseudo-digit so that they never share a match group with real keys
    const unsigned digit = ok ? (((uint32_t)k[r] >> shift) & 255u) : (256u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    int prev = 0;
    if (ok && lane == leader) {
      prev = wcount[warp][digit];
      wcount[warp][digit] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[r] = prev + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {  // warp offsets per digit: exclusive scan over the warps (thread = digit)
    int run = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) { const int c = wcount[w][threadIdx.x]; wcount[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int64_t i = first + r * 32 + lane;
    if (i < n) {
      const unsigned digit = ((uint32_t)k[r] >> shift) & 255u;
      const int64_t dst = (int64_t)dbase[digit] + wcount[warp][digit] + rank[r];
      keys_out[dst] = k[r];
      vals_out[dst] = v[r];
    }
  }
}

static inline int64_t sort_tiles(int64_t n) { return (n + kSortTile - 1) / kSortTile; }
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

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
                           int reso, int morton, int32_t* keys, int32_t* range_flag, t2h_stream_t stream) {
  if (!xyz || !keys || n_points < 0 || point_stride < 2 || n_per_batch <= 0 || reso <= 0 || reso > 32768)
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;  // Morton keys need a power of two
  int64_t n_batch = (n_points + n_per_batch - 1) / n_per_batch;
  if (n_batch * (int64_t)reso * reso > (int64_t)INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (n_points == 0) return T2H_OK;
  xy_keys_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n_points, point_stride, n_per_batch, reso, morton, keys, range_flag);
  T2H_CHECK_LAUNCH();
  return T2H_OK;
}

extern "C" int t2h_xy_keys_ragged(const float* xyz, int64_t n_points, int64_t point_stride, const int64_t* offsets,
                                  int n_tiles, int reso, int morton, int32_t* keys, int32_t* range_flag,
                                  t2h_stream_t stream) {
  if (!xyz || !keys || !offsets || n_points < 0 || point_stride < 2 || n_tiles <= 0 || reso <= 0 || reso > 32768)
    return T2H_ERR_INVALID_ARGUMENT;
  if (morton && (reso & (reso - 1))) return T2H_ERR_INVALID_ARGUMENT;
  if ((int64_t)n_tiles * reso * reso > (int64_t)INT32_MAX) return T2H_ERR_UNSUPPORTED_SHAPE;
  if (n_points == 0) return T2H_OK;
  xy_keys_ragged_kernel<<<blocks_for(n_points, 256), 256, 0, (cudaStream_t)stream>>>(xyz, n_points, point_stride, offsets,
                                                                                    n_tiles, reso, morton, keys, range_flag);
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
  // ping-pong (key, position) buffers + per-tile digit histograms + digit totals
  return 2 * align256((size_t)n_points * sizeof(int32_t)) + align256((size_t)256 * sort_tiles(n_points) * sizeof(int32_t)) +
         align256(256 * sizeof(int32_t)) + 256;
}

extern "C" int t2h_sort_by_cell(const int32_t* keys, int64_t n_points, int64_t n_keys, void* workspace,
                                size_t workspace_bytes, int32_t* keys_sorted, int32_t* perm,
                                int32_t* cell_start, t2h_stream_t stream) {
  if (!keys || !keys_sorted || !perm || !workspace || n_points < 0 || n_keys <= 0 ||
      n_points > (int64_t)INT32_MAX - 1 || n_keys > (int64_t)INT32_MAX - 1)
    return T2H_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < t2h_sort_workspace_bytes(n_points)) return T2H_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_points > 0) {
    const int64_t tiles = sort_tiles(n_points);
    char* ws = (char*)workspace;
    int32_t* alt_k = (int32_t*)ws;                     ws += align256((size_t)n_points * sizeof(int32_t));
    int32_t* alt_v = (int32_t*)ws;                     ws += align256((size_t)n_points * sizeof(int32_t));
    int32_t* hist = (int32_t*)ws;                      ws += align256((size_t)256 * tiles * sizeof(int32_t));
    int32_t* total = (int32_t*)ws;
    int bits = 1;
    while (bits < 31 && ((int64_t)1 << bits) < n_keys) ++bits;
    const int passes = (bits + 7) / 8;
    // ping-pong so that the LAST pass writes (keys_sorted, perm): an even number of passes starts via the alternates
    const int32_t* src_k = keys;
    const int32_t* src_v = nullptr;  // first pass: the value is the input position itself
    for (int pass = 0; pass < passes; ++pass) {
      const bool to_out = ((passes - 1 - pass) & 1) == 0;
      int32_t* dst_k = to_out ? keys_sorted : alt_k;
      int32_t* dst_v = to_out ? perm : alt_v;
      radix_hist_kernel<<<(unsigned)tiles, kSortThreads, 0, s>>>(src_k, n_points, 8 * pass, tiles, hist);
      T2H_CHECK_LAUNCH();
      radix_scan_kernel<<<256, kSortThreads, 0, s>>>(hist, tiles, total);
      T2H_CHECK_LAUNCH();
      radix_scatter_kernel<<<(unsigned)tiles, kSortThreads, 0, s>>>(src_k, src_v, n_points, 8 * pass, tiles, hist, total, dst_k, dst_v);
      T2H_CHECK_LAUNCH();
      src_k = dst_k;
      src_v = dst_v;
    }
  }
  if (cell_start) {  // nullable: a plain stable sort
    cell_start_kernel<<<blocks_for(n_points + 1, 256), 256, 0, s>>>(keys_sorted, n_points, n_keys, cell_start);
    T2H_CHECK_LAUNCH();
  }
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
