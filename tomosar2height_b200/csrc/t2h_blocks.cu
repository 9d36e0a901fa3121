// t2h_blocks.cu -- one-call entry points for the two point-MLP blocks of the path:
//   t2h_resblock_fwd / bwd   ResnetBlockFC   (reference block/resnet.py:46-54, instances pointnet.py:37-39,73-79)
//   t2h_comm_mlp_fwd / bwd   fc_comm + fc_c  (reference encoder/alto.py:63-69,123-128)
// They are COMPOSITIONS of the GEMM entry points of t2h_linear.cu, issued on the caller's stream in the order the
// Python mirror issues them (block/resnet.py, encoder/alto.py of this package) -- not single fused kernels; DESIGN.md
// §8.2 has the arithmetic that decided it.  What the composition fuses: ReLU-on-load, the concat halves as two K
// sources, bias, the shortcut / fc_c result as the residual of the last GEMM, the ReLU mask of the input gradient.
// Weights come as plain fp32 matrices; the operand splits (3xTF32 for narrow layers, 3xFP16 + power-of-two scales
// for n_out > 64 and K >= 128) and the transposes for the input gradients are produced in the caller's workspace.
// A workspace query runs the same plan without launching.
#include <cuda_runtime.h>
#include <stdint.h>

#include "t2h.h"
#include "t2h_common.cuh"

namespace t2h {
namespace blocks {

// ---- workspace arena: a dry run only counts -----------------------------------------------------------
struct Arena {
  char* base;
  size_t used, cap;
  bool dry;
  bool overflow;
  void* take(size_t bytes) {
    const size_t at = (used + 255) & ~(size_t)255;
    used = at + bytes;
    if (dry) return nullptr;
    if (used > cap) { overflow = true; return nullptr; }
    return base + at;
  }
};

struct Plan {
  Arena a;
  cudaStream_t s;
  int status;
  bool ok() const { return status == T2H_OK && !a.overflow; }
  void fail(int st) { if (status == T2H_OK) status = st; }
};

__global__ void transpose_kernel(const float* __restrict__ w, int64_t ld, int n, int k, float* __restrict__ wt) {
  // wt[c][r] = w[r][c] for r < n, c < k (weights: at most a few MB, once per call)
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < n && c < k) ? w[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < k && r < n) wt[(int64_t)c * n + r] = tile[threadIdx.x][i];
  }
}

static inline bool use_f16(int n_out, int k) { return n_out > 64 && k >= 128 && k % 8 == 0; }

// transpose of the columns [c0, c0 + k) of w [n][ld] into the arena: [k][n] (the weight of an input-gradient GEMM)
static const float* transposed(Plan& p, const float* w, int64_t ld, int n, int c0, int k) {
  float* dst = (float*)p.a.take((size_t)n * k * sizeof(float));
  if (p.a.dry || !p.ok()) return dst;
  transpose_kernel<<<dim3((k + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, p.s>>>(w + c0, ld, n, k, dst);
  if (cudaGetLastError() != cudaSuccess) p.fail(T2H_ERR_CUDA);
  return dst;
}

// out = act([x1 | x2]) @ w^T (+ bias) (* (mask > 0)) (+ residual);  w: dense fp32 [n_out][k1 + k2]
static void gemm(Plan& p, const float* x1, int64_t ld1, int k1, const float* x2, int64_t ld2, int k2, int64_t rows,
                 const float* w, int n_out, const float* bias, int relu_in, const float* mask, int64_t ld_mask,
                 const float* residual, int64_t ld_res, float* out, int64_t ld_out) {
  const int k = k1 + k2;
  const int64_t n = (int64_t)n_out * k;
  if (use_f16(n_out, k)) {
    uint32_t* slots = (uint32_t*)p.a.take(2 * sizeof(uint32_t));
    uint16_t* hi = (uint16_t*)p.a.take((size_t)n * 2);
    uint16_t* lo = (uint16_t*)p.a.take((size_t)n * 2);
    if (p.a.dry || !p.ok()) return;
    int st = t2h_absmax(x1, ld1, k1, x2, ld2, k2, rows, slots, p.s);
    if (!st) st = t2h_absmax(w, k, k, nullptr, 0, 0, n_out, slots + 1, p.s);
    if (!st) st = t2h_split_f16(w, n, slots + 1, hi, lo, p.s);
    if (!st) st = t2h_linear_fwd_f16(x1, ld1, k1, x2, ld2, k2, rows, slots, hi, lo, slots + 1, n_out, bias, relu_in, mask,
                                     ld_mask, residual, ld_res, out, ld_out, nullptr, p.s);
    p.fail(st);
    return;
  }
  float* hi = (float*)p.a.take((size_t)n * sizeof(float));
  float* lo = (float*)p.a.take((size_t)n * sizeof(float));
  if (p.a.dry || !p.ok()) return;
  int st = t2h_split_tf32(w, n, hi, lo, p.s);
  if (!st) st = t2h_linear_fwd(x1, ld1, k1, x2, ld2, k2, rows, hi, lo, n_out, bias, relu_in, mask, ld_mask, residual, ld_res,
                               out, ld_out, p.s);
  p.fail(st);
}

// d_w[n_out][k_in] (pitch ld_w) = g^T act(x), d_b (nullable) = column sums of g
static void wgrad(Plan& p, const float* g, int64_t ld_g, const float* x, int64_t ld_x, int64_t rows, int n_out, int k_in,
                  int relu_in, float* d_w, int64_t ld_w, float* d_b) {
  const size_t ws_bytes = t2h_linear_wgrad_workspace_bytes(rows, n_out, k_in);
  void* ws = p.a.take(ws_bytes);
  const bool f16 = k_in >= 128 && n_out >= 32;
  uint32_t* slots = f16 ? (uint32_t*)p.a.take(2 * sizeof(uint32_t)) : nullptr;
  if (p.a.dry || !p.ok()) return;
  int st;
  if (f16) {
    st = t2h_absmax(g, ld_g, n_out, nullptr, 0, 0, rows, slots, p.s);
    if (!st) st = t2h_absmax(x, ld_x, k_in, nullptr, 0, 0, rows, slots + 1, p.s);
    if (!st) st = t2h_linear_wgrad_f16(g, ld_g, slots, x, ld_x, slots + 1, rows, n_out, k_in, relu_in, ws, ws_bytes, d_w, ld_w,
                                       d_b, p.s);
  } else {
    st = t2h_linear_wgrad(g, ld_g, x, ld_x, rows, n_out, k_in, relu_in, ws, ws_bytes, d_w, ld_w, d_b, p.s);
  }
  p.fail(st);
}

// ---- ResnetBlockFC -----------------------------------------------------------------------------------
struct ResArgs {
  const float *x1, *x2;
  int64_t ld_x1, ld_x2, rows;
  int k1, k2, n_h, n_out;
  const float *w0, *b0, *w1, *b1, *ws;
};

static bool res_shapes_ok(const ResArgs& r) {
  if (r.rows < 0 || r.k1 <= 0 || r.k2 < 0 || r.n_h <= 0 || r.n_out <= 0) return false;
  if ((r.k1 % 4) || (r.k2 % 4) || (r.n_h % 4) || (r.n_out % 4)) return false;
  if (r.k2 && (r.k1 % 32)) return false;  // two K sources: the first ends on a 32-wide chunk
  if (!r.ws && (r.k2 != 0 || r.k1 != r.n_out)) return false;  // identity shortcut: size_in == size_out, one source
  return true;
}

static void resblock_fwd_plan(Plan& p, const ResArgs& r, float* net, int64_t ld_net, float* out, int64_t ld_out) {
  gemm(p, r.x1, r.ld_x1, r.k1, r.x2, r.ld_x2, r.k2, r.rows, r.w0, r.n_h, r.b0, 1, nullptr, 0, nullptr, 0, net, ld_net);
  if (r.ws) {
    gemm(p, r.x1, r.ld_x1, r.k1, r.x2, r.ld_x2, r.k2, r.rows, r.ws, r.n_out, nullptr, 0, nullptr, 0, nullptr, 0, out, ld_out);
    gemm(p, net, ld_net, r.n_h, nullptr, 0, 0, r.rows, r.w1, r.n_out, r.b1, 1, nullptr, 0, out, ld_out, out, ld_out);
  } else {
    gemm(p, net, ld_net, r.n_h, nullptr, 0, 0, r.rows, r.w1, r.n_out, r.b1, 1, nullptr, 0, r.x1, r.ld_x1, out, ld_out);
  }
}

struct ResGrads {
  float *d_x1, *d_x2;
  int64_t ld_dx1, ld_dx2;
  float *d_w0, *d_b0, *d_w1, *d_b1, *d_ws;
};

static void resblock_bwd_plan(Plan& p, const ResArgs& r, const float* g, int64_t ld_g, const float* net, int64_t ld_net,
                              const ResGrads& o) {
  const int n_in = r.k1 + r.k2;
  // fc_1: weight / bias gradient, then the gradient of its (pre-activation) input
  wgrad(p, g, ld_g, net, ld_net, r.rows, r.n_out, r.n_h, 1, o.d_w1, r.n_h, o.d_b1);
  float* g_net = (float*)p.a.take((size_t)(r.rows > 0 ? r.rows : 1) * r.n_h * sizeof(float));
  const float* w1t = transposed(p, r.w1, r.n_h, r.n_out, 0, r.n_h);  // [n_h][n_out]
  gemm(p, g, ld_g, r.n_out, nullptr, 0, 0, r.rows, w1t, r.n_h, nullptr, 0, net, ld_net, nullptr, 0, g_net, r.n_h);
  // fc_0 and the shortcut, source by source
  const float* xs[2] = {r.x1, r.x2};
  const int64_t lds[2] = {r.ld_x1, r.ld_x2};
  const int ks[2] = {r.k1, r.k2};
  float* dxs[2] = {o.d_x1, o.d_x2};
  const int64_t ld_dx[2] = {o.ld_dx1, o.ld_dx2};
  int c0 = 0;
  for (int i = 0; i < 2; ++i) {
    if (ks[i] == 0) continue;
    wgrad(p, g_net, r.n_h, xs[i], lds[i], r.rows, r.n_h, ks[i], 1, o.d_w0 + c0, n_in, i == 0 ? o.d_b0 : nullptr);
    if (r.ws) wgrad(p, g, ld_g, xs[i], lds[i], r.rows, r.n_out, ks[i], 0, o.d_ws + c0, n_in, nullptr);
    if (dxs[i]) {
      const float* w0t = transposed(p, r.w0, n_in, r.n_h, c0, ks[i]);  // [k_i][n_h]
      if (r.ws) {
        const float* wst = transposed(p, r.ws, n_in, r.n_out, c0, ks[i]);  // [k_i][n_out]
        gemm(p, g, ld_g, r.n_out, nullptr, 0, 0, r.rows, wst, ks[i], nullptr, 0, nullptr, 0, nullptr, 0, dxs[i], ld_dx[i]);
        gemm(p, g_net, r.n_h, r.n_h, nullptr, 0, 0, r.rows, w0t, ks[i], nullptr, 0, xs[i], lds[i], dxs[i], ld_dx[i], dxs[i], ld_dx[i]);
      } else {
        gemm(p, g_net, r.n_h, r.n_h, nullptr, 0, 0, r.rows, w0t, ks[i], nullptr, 0, xs[i], lds[i], g, ld_g, dxs[i], ld_dx[i]);
      }
    }
    c0 += ks[i];
  }
}

// ---- fc_comm + fc_c -----------------------------------------------------------------------------------
struct CommArgs {
  const float *c, *c_last;
  int64_t ld_c, ld_cl, rows;
  int C, C_prev;
  const float *w0, *b0, *w2, *b2, *wc, *bc;
};

static bool comm_shapes_ok(const CommArgs& a) {
  if (a.rows < 0 || a.C <= 0 || (a.C % 4)) return false;
  if (a.c_last && (a.C_prev <= 0 || (a.C_prev % 4) || !a.wc)) return false;
  return true;
}

static void comm_fwd_plan(Plan& p, const CommArgs& a, float* hidden, int64_t ld_h, float* out, int64_t ld_out) {
  gemm(p, a.c, a.ld_c, a.C, nullptr, 0, 0, a.rows, a.w0, 2 * a.C, a.b0, 0, nullptr, 0, nullptr, 0, hidden, ld_h);
  if (a.c_last) {
    gemm(p, a.c_last, a.ld_cl, a.C_prev, nullptr, 0, 0, a.rows, a.wc, a.C, a.bc, 0, nullptr, 0, nullptr, 0, out, ld_out);
    gemm(p, hidden, ld_h, 2 * a.C, nullptr, 0, 0, a.rows, a.w2, a.C, a.b2, 1, nullptr, 0, out, ld_out, out, ld_out);
  } else {
    gemm(p, hidden, ld_h, 2 * a.C, nullptr, 0, 0, a.rows, a.w2, a.C, a.b2, 1, nullptr, 0, nullptr, 0, out, ld_out);
  }
}

struct CommGrads {
  float *d_c, *d_c_last;
  int64_t ld_dc, ld_dcl;
  float *d_w0, *d_b0, *d_w2, *d_b2, *d_wc, *d_bc;
};

static void comm_bwd_plan(Plan& p, const CommArgs& a, const float* g, int64_t ld_g, const float* hidden, int64_t ld_h,
                          const CommGrads& o) {
  const int H = 2 * a.C;
  wgrad(p, g, ld_g, hidden, ld_h, a.rows, a.C, H, 1, o.d_w2, H, o.d_b2);
  float* g_h = (float*)p.a.take((size_t)(a.rows > 0 ? a.rows : 1) * H * sizeof(float));
  const float* w2t = transposed(p, a.w2, H, a.C, 0, H);  // [2C][C]
  gemm(p, g, ld_g, a.C, nullptr, 0, 0, a.rows, w2t, H, nullptr, 0, hidden, ld_h, nullptr, 0, g_h, H);
  wgrad(p, g_h, H, a.c, a.ld_c, a.rows, H, a.C, 0, o.d_w0, a.C, o.d_b0);
  if (o.d_c) {
    const float* w0t = transposed(p, a.w0, a.C, H, 0, a.C);  // [C][2C]
    gemm(p, g_h, H, H, nullptr, 0, 0, a.rows, w0t, a.C, nullptr, 0, nullptr, 0, nullptr, 0, o.d_c, o.ld_dc);
  }
  if (a.c_last) {
    wgrad(p, g, ld_g, a.c_last, a.ld_cl, a.rows, a.C, a.C_prev, 0, o.d_wc, a.C_prev, o.d_bc);
    if (o.d_c_last) {
      const float* wct = transposed(p, a.wc, a.C_prev, a.C, 0, a.C_prev);  // [C_prev][C]
      gemm(p, g, ld_g, a.C, nullptr, 0, 0, a.rows, wct, a.C_prev, nullptr, 0, nullptr, 0, nullptr, 0, o.d_c_last, o.ld_dcl);
    }
  }
}

static Plan make_plan(void* ws, size_t bytes, cudaStream_t s, bool dry) {
  Plan p;
  p.a.base = (char*)ws; p.a.used = 0; p.a.cap = bytes; p.a.dry = dry; p.a.overflow = false;
  p.s = s; p.status = T2H_OK;
  return p;
}

static int finish(const Plan& p) {
  if (p.a.overflow) return T2H_ERR_WORKSPACE_TOO_SMALL;
  return p.status;
}

}  // namespace blocks
}  // namespace t2h

using namespace t2h::blocks;

extern "C" size_t t2h_resblock_workspace_bytes(int64_t rows, int k1, int k2, int n_h, int n_out, int has_shortcut) {
  ResArgs r = {};
  r.rows = rows; r.k1 = k1; r.k2 = k2; r.n_h = n_h; r.n_out = n_out;
  r.ld_x1 = k1; r.ld_x2 = k2;
  r.ws = has_shortcut ? (const float*)16 : nullptr;  // only tested for null in a dry run
  if (!res_shapes_ok(r)) return 256;
  Plan f = make_plan(nullptr, 0, nullptr, true);
  resblock_fwd_plan(f, r, nullptr, n_h, nullptr, n_out);
  Plan b = make_plan(nullptr, 0, nullptr, true);
  ResGrads o = {};
  o.d_x1 = (float*)16; o.d_x2 = k2 ? (float*)16 : nullptr;
  resblock_bwd_plan(b, r, nullptr, n_out, nullptr, n_h, o);
  return (f.a.used > b.a.used ? f.a.used : b.a.used) + 256;
}

extern "C" int t2h_resblock_fwd(const float* x1, int64_t ld_x1, int k1, const float* x2, int64_t ld_x2, int k2, int64_t rows,
                                const float* w0, const float* b0, int n_h, const float* w1, const float* b1,
                                const float* w_shortcut, int n_out, void* workspace, size_t workspace_bytes, float* net,
                                int64_t ld_net, float* out, int64_t ld_out, t2h_stream_t stream) {
  ResArgs r = {x1, x2, ld_x1, ld_x2, rows, k1, x2 ? k2 : 0, n_h, n_out, w0, b0, w1, b1, w_shortcut};
  if (!x1 || !w0 || !w1 || !net || !out || !workspace || !res_shapes_ok(r)) return T2H_ERR_INVALID_ARGUMENT;
  if (rows == 0) return T2H_OK;
  Plan p = make_plan(workspace, workspace_bytes, (cudaStream_t)stream, false);
  resblock_fwd_plan(p, r, net, ld_net, out, ld_out);
  return finish(p);
}

extern "C" int t2h_resblock_bwd(const float* grad_out, int64_t ld_g, const float* x1, int64_t ld_x1, int k1, const float* x2,
                                int64_t ld_x2, int k2, const float* net, int64_t ld_net, int64_t rows, const float* w0,
                                int n_h, const float* w1, const float* w_shortcut, int n_out, void* workspace,
                                size_t workspace_bytes, float* d_x1, int64_t ld_dx1, float* d_x2, int64_t ld_dx2, float* d_w0,
                                float* d_b0, float* d_w1, float* d_b1, float* d_w_shortcut, t2h_stream_t stream) {
  ResArgs r = {x1, x2, ld_x1, ld_x2, rows, k1, x2 ? k2 : 0, n_h, n_out, w0, nullptr, w1, nullptr, w_shortcut};
  if (!grad_out || !x1 || !net || !w0 || !w1 || !workspace || !d_w0 || !d_w1 || !res_shapes_ok(r)) return T2H_ERR_INVALID_ARGUMENT;
  if (w_shortcut && !d_w_shortcut) return T2H_ERR_INVALID_ARGUMENT;
  ResGrads o = {d_x1, r.k2 ? d_x2 : nullptr, ld_dx1, ld_dx2, d_w0, d_b0, d_w1, d_b1, d_w_shortcut};
  Plan p = make_plan(workspace, workspace_bytes, (cudaStream_t)stream, false);
  resblock_bwd_plan(p, r, grad_out, ld_g, net, ld_net, o);
  return finish(p);
}

extern "C" size_t t2h_comm_mlp_workspace_bytes(int64_t rows, int C, int C_prev) {
  CommArgs a = {};
  a.rows = rows; a.C = C; a.C_prev = C_prev; a.ld_c = C; a.ld_cl = C_prev;
  a.c_last = C_prev > 0 ? (const float*)16 : nullptr;
  a.wc = a.c_last;
  if (!comm_shapes_ok(a)) return 256;
  Plan f = make_plan(nullptr, 0, nullptr, true);
  comm_fwd_plan(f, a, nullptr, 2 * C, nullptr, C);
  Plan b = make_plan(nullptr, 0, nullptr, true);
  CommGrads o = {};
  o.d_c = (float*)16; o.d_c_last = a.c_last ? (float*)16 : nullptr;
  comm_bwd_plan(b, a, nullptr, C, nullptr, 2 * C, o);
  return (f.a.used > b.a.used ? f.a.used : b.a.used) + 256;
}

extern "C" int t2h_comm_mlp_fwd(const float* c, int64_t ld_c, int C, const float* c_last, int64_t ld_cl, int C_prev,
                                int64_t rows, const float* w0, const float* b0, const float* w2, const float* b2,
                                const float* wc, const float* bc, void* workspace, size_t workspace_bytes, float* hidden,
                                int64_t ld_hidden, float* out, int64_t ld_out, t2h_stream_t stream) {
  CommArgs a = {c, c_last, ld_c, ld_cl, rows, C, c_last ? C_prev : 0, w0, b0, w2, b2, wc, bc};
  if (!c || !w0 || !w2 || !hidden || !out || !workspace || !comm_shapes_ok(a)) return T2H_ERR_INVALID_ARGUMENT;
  if (rows == 0) return T2H_OK;
  Plan p = make_plan(workspace, workspace_bytes, (cudaStream_t)stream, false);
  comm_fwd_plan(p, a, hidden, ld_hidden, out, ld_out);
  return finish(p);
}

extern "C" int t2h_comm_mlp_bwd(const float* grad_out, int64_t ld_g, const float* c, int64_t ld_c, int C, const float* c_last,
                                int64_t ld_cl, int C_prev, const float* hidden, int64_t ld_hidden, int64_t rows,
                                const float* w0, const float* w2, const float* wc, void* workspace, size_t workspace_bytes,
                                float* d_c, int64_t ld_dc, float* d_c_last, int64_t ld_dcl, float* d_w0, float* d_b0,
                                float* d_w2, float* d_b2, float* d_wc, float* d_bc, t2h_stream_t stream) {
  CommArgs a = {c, c_last, ld_c, ld_cl, rows, C, c_last ? C_prev : 0, w0, nullptr, w2, nullptr, wc, nullptr};
  if (!grad_out || !c || !hidden || !w0 || !w2 || !workspace || !d_w0 || !d_w2 || !comm_shapes_ok(a)) return T2H_ERR_INVALID_ARGUMENT;
  if (c_last && !d_wc) return T2H_ERR_INVALID_ARGUMENT;
  CommGrads o = {d_c, c_last ? d_c_last : nullptr, ld_dc, ld_dcl, d_w0, d_b0, d_w2, d_b2, d_wc, d_bc};
  Plan p = make_plan(workspace, workspace_bytes, (cudaStream_t)stream, false);
  comm_bwd_plan(p, a, grad_out, ld_g, hidden, ld_hidden, o);
  return finish(p);
}
