"""nn.Linear on the point path as tcgen05 3xTF32 GEMMs (t2h_linear_fwd / t2h_linear_wgrad / t2h_colsum).

``linear(x, weight, bias, x2=None, relu_in=False, residual=None)`` computes

    relu?( [x | x2] ) @ weight.T + bias + residual

with fp32 inputs / outputs and fp32-grade accuracy, and is differentiable: the backward runs the
input-gradient GEMM through the same forward kernel (weight transposed, ReLU mask and gradient
accumulation fused in the epilogue) and the weight gradient through the MN-major kernel.  Weights stay
ordinary fp32 ``nn.Parameter``s; their TF32 hi/lo splits (and transposes) are derived caches keyed on
the parameter's version counter.
"""
import os
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import ptr


class _SplitCache:
    """(weight, column slice, transposed?) -> (hi, lo), invalidated when the parameter changes."""

    def __init__(self):
        self._store = {}

    def get(self, weight, c0=0, c1=None, transposed=False):
        c1 = weight.shape[1] if c1 is None else c1
        key = (id(weight), c0, c1, transposed)
        hit = None if capture_mode else self._store.get(key)
        if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
            return hit[3], hit[4]
        with torch.no_grad():
            w = weight.detach()[:, c0:c1]
            w = w.t().contiguous() if transposed else w.contiguous()
            hi, lo = torch.empty_like(w), torch.empty_like(w)
            _lib.call("t2h_split_tf32", ptr(w), w.numel(), ptr(hi), ptr(lo))
        if capture_mode:
            return hi, lo  # lives in the graph's private pool; not a cache entry
        if len(self._store) > 4096:
            self._store.clear()
        self._store[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), hi, lo)
        return hi, lo

    def get_matrix(self, weight, tag, builder):
        """(hi, lo) split of ``builder(weight.detach())`` (a 2-D fp32 matrix), cached per parameter version."""
        key = (id(weight), tag)
        hit = None if capture_mode else self._store.get(key)
        if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
            return hit[3], hit[4]
        with torch.no_grad():
            w = builder(weight.detach()).contiguous()
            hi, lo = torch.empty_like(w), torch.empty_like(w)
            _lib.call("t2h_split_tf32", ptr(w), w.numel(), ptr(hi), ptr(lo))
        if capture_mode:
            return hi, lo
        if len(self._store) > 4096:
            self._store.clear()
        self._store[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), hi, lo)
        return hi, lo


_cache = _SplitCache()
# True while a CUDA graph is being captured (graph.py): splits are recomputed inside the graph, because the
# weights change between replays
capture_mode = False
# ablation switch for benchmarks / debugging only: run the point MLPs as plain cuBLAS fp32 GEMMs
USE_LIBRARY_GEMM = os.environ.get("T2H_LINEAR", "") == "cublas"


def _rowmajor(t):
    """2-D fp32 CUDA view usable by TMA: unit column stride, 16-byte aligned base and pitch."""
    if t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def _launch_fwd(x1, x2, w_hi, w_lo, n_out, bias, relu_in, mask, residual, out):
    rows, k1 = x1.shape
    k2 = 0 if x2 is None else x2.shape[1]
    _lib.call("t2h_linear_fwd", ptr(x1), x1.stride(0), k1, ptr(x2), 0 if x2 is None else x2.stride(0), k2, rows,
              ptr(w_hi), ptr(w_lo), n_out, ptr(bias), int(relu_in), ptr(mask), 0 if mask is None else mask.stride(0),
              ptr(residual), 0 if residual is None else residual.stride(0), ptr(out), out.stride(0))


def _launch_wgrad(gy, x, relu_in, grad_w_view, grad_b=None):
    rows, n_out = gy.shape
    k_in = x.shape[1]
    lib = _lib.load()
    ws_bytes = int(lib.t2h_linear_wgrad_workspace_bytes(rows, n_out, k_in))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gy.device)
    _lib.call("t2h_linear_wgrad", ptr(gy), gy.stride(0), ptr(x), x.stride(0), rows, n_out, k_in, int(relu_in), ptr(ws),
              ws_bytes, ptr(grad_w_view), grad_w_view.stride(0), ptr(grad_b))


def colsum(g):
    rows, n = g.shape
    lib = _lib.load()
    ws_bytes = int(lib.t2h_colsum_workspace_bytes(rows, n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
    out = torch.empty(n, dtype=torch.float32, device=g.device)
    _lib.call("t2h_colsum", ptr(g), g.stride(0), rows, n, ptr(ws), ws_bytes, ptr(out))
    return out


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, weight, bias, residual, relu_in):
        x1 = _rowmajor(x1)
        x2 = None if x2 is None else _rowmajor(x2)
        residual = None if residual is None else _rowmajor(residual)
        n_out = weight.shape[0]
        w_hi, w_lo = _cache.get(weight)
        out = torch.empty(x1.shape[0], n_out, dtype=torch.float32, device=x1.device)
        _launch_fwd(x1, x2, w_hi, w_lo, n_out, bias, relu_in, None, residual, out)
        ctx.save_for_backward(x1, x2, weight)
        ctx.relu_in = relu_in
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, gy):
        x1, x2, weight = ctx.saved_tensors
        gy = _rowmajor(gy)
        k1 = x1.shape[1]
        n_out, k_total = weight.shape
        need = ctx.needs_input_grad
        d_x1 = d_x2 = d_w = d_b = d_res = None
        if need[0]:
            t_hi, t_lo = _cache.get(weight, 0, k1, transposed=True)
            d_x1 = torch.empty_like(x1)
            _launch_fwd(gy, None, t_hi, t_lo, k1, None, False, x1 if ctx.relu_in else None, None, d_x1)
        if x2 is not None and need[1]:
            t_hi, t_lo = _cache.get(weight, k1, k_total, transposed=True)
            d_x2 = torch.empty_like(x2)
            _launch_fwd(gy, None, t_hi, t_lo, k_total - k1, None, False, x2 if ctx.relu_in else None, None, d_x2)
        want_bias = ctx.has_bias and need[3]
        if need[2]:
            d_w = torch.empty(n_out, k_total, dtype=torch.float32, device=gy.device)
            if want_bias:  # the bias gradient (column sum of gy) rides along with the first wgrad launch
                d_b = torch.empty(n_out, dtype=torch.float32, device=gy.device)
            _launch_wgrad(gy, x1, ctx.relu_in, d_w[:, :k1], d_b)
            if x2 is not None:
                _launch_wgrad(gy, x2, ctx.relu_in, d_w[:, k1:])
        elif want_bias:
            d_b = colsum(gy)
        if need[4]:
            d_res = gy
        return d_x1, d_x2, d_w, d_b, d_res, None


def _pad_to(v, m):
    return (m - v % m) % m


def linear(x, weight, bias=None, x2=None, relu_in=False, residual=None):
    """relu?([x | x2]) @ weight.T + bias + residual over the last dimension (any leading shape).

    Widths that TMA cannot address directly (rows must be multiples of 16 bytes) are zero-padded to a
    multiple of 4 -- e.g. the 1-wide ``fc_out`` head of the FC decoder (pixel.py:51) -- and the result
    is sliced back; padding is differentiable, so gradients reach the original parameters.
    """
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("linear: expected a float32 CUDA tensor (the B200 path has no CPU fallback)")
    if USE_LIBRARY_GEMM:  # ablation only (T2H_LINEAR=cublas)
        xin = x if x2 is None else torch.cat([x, x2], dim=-1)
        y = F.linear(F.relu(xin) if relu_in else xin, weight, bias)
        return y if residual is None else y + residual
    n_out, k1 = weight.shape[0], x.shape[-1]
    if x2 is not None and (k1 % 32 or x2.shape[-1] % 4):
        x, x2, k1 = torch.cat([x, x2], dim=-1), None, k1 + x2.shape[-1]  # second source must start on a 32-wide chunk
    pk, pn = _pad_to(k1, 4) if x2 is None else 0, _pad_to(n_out, 4)
    if pk:
        x = F.pad(x, (0, pk))
        weight = F.pad(weight, (0, pk))
    if pn:
        weight = F.pad(weight, (0, 0, 0, pn))
        bias = None if bias is None else F.pad(bias, (0, pn))
        residual = None if residual is None else F.pad(residual, (0, pn))
    lead = x.shape[:-1]
    x2d = x.reshape(-1, x.shape[-1])
    x22d = None if x2 is None else x2.reshape(-1, x2.shape[-1])
    res2d = None if residual is None else residual.reshape(-1, weight.shape[0])
    y = _LinearTC.apply(x2d, x22d, weight, bias, res2d, bool(relu_in))
    y = y.view(*lead, weight.shape[0])
    return y[..., :n_out] if pn else y
