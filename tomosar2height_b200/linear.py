"""nn.Linear on the point path as tcgen05 split-precision GEMMs (t2h_linear_fwd[_f16] / t2h_linear_wgrad[_f16] /
t2h_colsum): 3xTF32 for narrow layers, 3xFP16 with power-of-two operand scaling for K >= 128.

``linear(x, weight, bias, x2=None, relu_in=False, residual=None)`` computes

    relu?( [x | x2] ) @ weight.T + bias + residual

with fp32 inputs / outputs and fp32-grade accuracy, and is differentiable: the backward runs the
input-gradient GEMM through the same forward kernel (weight transposed, ReLU mask and gradient
accumulation fused in the epilogue) and the weight gradient through the MN-major kernel.  Weights stay
ordinary fp32 ``nn.Parameter``s; their hi/lo splits (and transposes) are derived caches keyed on
the parameter's version counter.  The fp16 flavour needs the maximum magnitude of every operand: the
GEMM epilogues publish it for their outputs, ``_AbsmaxRegistry`` hands it to the consumer, and only tensors
that come from elsewhere take a streaming ``t2h_absmax`` pass.
"""
import os
import weakref

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib
from ._lib import ptr


class _SplitCache:
    """(weight, column slice / tag, flavour) -> operand split, invalidated when the parameter changes.

    tf32 flavour: (hi, lo) fp32 matrices; f16 flavour: (hi, lo, absmax) with fp16 matrices and the device
    word holding the bit pattern of max |w| that fixes the power-of-two scale (t2h_absmax / t2h_split_f16)."""

    def __init__(self):
        self._store = {}

    @staticmethod
    def _buffers(shape, device, f16):
        if not f16:
            return torch.empty(shape, dtype=torch.float32, device=device), torch.empty(shape, dtype=torch.float32, device=device)
        return (torch.empty(shape, dtype=torch.float16, device=device), torch.empty(shape, dtype=torch.float16, device=device),
                torch.empty(1, dtype=torch.int32, device=device))

    @staticmethod
    def _split_into(w, f16, bufs):
        """split of the dense 2-D fp32 matrix ``w`` into existing buffers (``_buffers``)"""
        if not f16:
            _lib.call("t2h_split_tf32", ptr(w), w.numel(), ptr(bufs[0]), ptr(bufs[1]))
        else:
            _lib.call("t2h_absmax", ptr(w), w.shape[1], w.shape[1], None, 0, 0, w.shape[0], ptr(bufs[2]))
            _lib.call("t2h_split_f16", ptr(w), w.numel(), ptr(bufs[2]), ptr(bufs[0]), ptr(bufs[1]))
        return bufs

    @classmethod
    def _split(cls, w, f16):
        return cls._split_into(w, f16, cls._buffers(w.shape, w.device, f16))

    def _lookup(self, weight, key, builder, f16):
        key = key + (f16,)
        if static_params is not None:
            # graph.py, eager warm-up before a capture and the capture itself.  A split of a PARAMETER does not belong
            # into the graph (it would be recomputed by every replay although the weights only change once per
            # optimizer step): the warm-up gives it persistent buffers -- allocated OUTSIDE the graph's memory pool,
            # whose blocks are recycled between the captured kernels -- and a recipe; the capture only hands the
            # buffers out, and GraphedTrainStep rewrites them eagerly whenever a parameter's version has moved.
            param = static_params.get(weight.data_ptr())
            if param is not None and param.shape == weight.shape:
                skey = (id(param),) + key[1:]
                rec = static_recipes.get(skey)
                if rec is None and not capture_mode:
                    with torch.no_grad():
                        shape = builder(param.detach()).shape
                    rec = static_recipes[skey] = (param, builder, f16, self._buffers(shape, param.device, f16))
                if rec is not None:
                    if not capture_mode:
                        with torch.no_grad():
                            self._split_into(builder(param.detach()).contiguous(), f16, rec[3])
                    return rec[3]
        if capture_mode:
            # splits of temporaries (padded / sliced weights, produced by captured kernels) stay inside the graph
            with torch.no_grad():
                return self._split(builder(weight.detach()).contiguous(), f16)  # lives in the graph's private pool
        hit = self._store.get(key)
        if hit is not None and hit[0]() is weight and hit[1] == weight._version and hit[2] == weight.data_ptr():
            return hit[3]
        with torch.no_grad():
            split = self._split(builder(weight.detach()).contiguous(), f16)
        if len(self._store) > 256:
            # entries of temporaries (reshaped / padded / permuted weight views are new tensors on every call) only
            # hold dead weak references: drop them so that their hi / lo device tensors are freed
            self._store = {k: v for k, v in self._store.items() if v[0]() is not None}
            if len(self._store) > 4096:
                self._store.clear()
        self._store[key] = (weakref.ref(weight), weight._version, weight.data_ptr(), split)
        return split

    def get(self, weight, c0=0, c1=None, transposed=False, f16=False):
        c1 = weight.shape[1] if c1 is None else c1
        return self._lookup(weight, (id(weight), c0, c1, transposed),
                            lambda w: w[:, c0:c1].t() if transposed else w[:, c0:c1], f16)

    def invalidate(self):
        """forget every derived split (call after the weights were changed behind autograd's back, e.g. by a captured
        optimizer step or a ``.data`` write, which do not bump the version counter)"""
        self._store.clear()

    def get_matrix(self, weight, tag, builder, f16=False):
        """split of ``builder(weight.detach())`` (a 2-D fp32 matrix), cached per parameter version."""
        return self._lookup(weight, (id(weight), tag), builder, f16)


_cache = _SplitCache()
# True while a CUDA graph is being captured (graph.py).  The weights change between replays: splits of parameters
# live in persistent buffers that graph.py refreshes (static_params: data_ptr -> Parameter, static_recipes: the
# recipes collected during the capture); any other split is recomputed inside the graph
capture_mode = False
static_params = None
static_recipes = None


def refresh_static_splits(recipes):
    """recompute every hoisted parameter split into its persistent buffers (eager launches on the current stream)"""
    with torch.no_grad():
        for param, builder, f16, bufs in recipes.values():
            _SplitCache._split_into(builder(param.detach()).contiguous(), f16, bufs)
# ablation switch for benchmarks / debugging only: run the point MLPs as plain cuBLAS fp32 GEMMs
USE_LIBRARY_GEMM = os.environ.get("T2H_LINEAR", "") == "cublas"
# wide layers run the 3xFP16 flavour (twice the tensor-core rate of 3xTF32, same fp32-grade accuracy);
# T2H_LINEAR_F16=0 keeps everything on 3xTF32 (ablation)
USE_F16 = os.environ.get("T2H_LINEAR_F16", "1") != "0"
USE_F16_WGRAD = USE_F16 and os.environ.get("T2H_WGRAD_F16", "1") != "0"


def use_f16(n_out, k_total):
    return USE_F16 and n_out > 64 and k_total >= 128 and k_total % 8 == 0


def _dense(t):
    """the tensor covers exactly ``numel`` consecutive elements from its data pointer (row-major, or a channels-last
    NCHW view of a row-major NHWC plane): its maximum is a property of that memory range, whatever the view"""
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


class _AbsmaxRegistry:
    """Device words holding max |t| of live tensors, so that an operand maximum is computed once.

    The fp16 GEMM publishes the maximum of its OUTPUT from the epilogue; a standalone ``t2h_absmax`` pass
    registers its result too (the same gradient often feeds two GEMMs).  An entry is keyed on the data
    pointer and only valid while the registered tensor object is alive (its memory cannot have been handed
    to another tensor) and unmodified (same version counter); anything else is a miss and falls back to a
    fresh pass, so a stale entry can never be used."""

    def __init__(self):
        self._entries = {}

    def put(self, t, slot):
        key = (t.data_ptr(), t.numel())
        entries = self._entries

        def drop(ref, key=key):
            if entries.get(key, (None,))[0] is ref:
                del entries[key]

        if t.is_inference():  # no version counter: cannot tell whether it was modified
            return
        entries[key] = (weakref.ref(t, drop), t._version, slot)

    def get(self, t):
        e = self._entries.get((t.data_ptr(), t.numel()))
        if e is None or not _dense(t) or t.is_inference():
            return None
        owner = e[0]()
        if owner is None or owner._version != e[1] or t._version != e[1] or not _dense(owner):
            return None
        return e[2]


_absmax = _AbsmaxRegistry()


def operand_absmax(x1, x2=None, owner=None):
    """device word with max |[x1 | x2]| (registered result if there is one, else one streaming pass).
    ``owner``: the tensor object whose lifetime guards the registry entry when ``x1`` is a temporary 2-D view."""
    owner = x1 if owner is None else owner
    if x2 is None:
        slot = _absmax.get(owner)
        if slot is not None:
            return slot
    slot = torch.empty(1, dtype=torch.int32, device=x1.device)
    _lib.call("t2h_absmax", ptr(x1), x1.stride(0), x1.shape[1], ptr(x2), 0 if x2 is None else x2.stride(0),
              0 if x2 is None else x2.shape[1], x1.shape[0], ptr(slot))
    if x2 is None and _dense(owner):
        _absmax.put(owner, slot)
    return slot


def publish_absmax(t, slot):
    """register ``slot`` (written by a kernel epilogue) as the maximum of the dense tensor ``t``."""
    if _dense(t):
        _absmax.put(t, slot)


class _Fork2(torch.autograd.Function):
    """Identity with two outputs.  A tensor that feeds two consumers gets its gradient as the sum of two
    branches; autograd would add them with a library kernel and the fp16 GEMM that consumes the sum would then
    need a pass of its own for the operand maximum.  The backward of this node does both in one kernel."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None or g2 is None:
            return g1 if g2 is None else g2
        if not (g1.is_cuda and g1.dtype == torch.float32 and g1.is_contiguous() and g2.is_contiguous()
                and g1.numel() % 4 == 0 and g1.data_ptr() % 16 == 0 and g2.data_ptr() % 16 == 0):
            return g1 + g2
        out = torch.empty_like(g1)
        slot = torch.empty(1, dtype=torch.int32, device=g1.device)
        _lib.call("t2h_add_absmax", ptr(g1), ptr(g2), g1.numel(), ptr(out), ptr(slot))
        publish_absmax(out, slot)
        return out


def fork2(x):
    """``x`` twice, for two consumers (see ``_Fork2``); a plain pass-through when no gradient is needed."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return x, x
    return _Fork2.apply(x)


def relu(x, slope=0.0):
    """F.relu / F.leaky_relu that hands the operand maximum on: max |relu(x)| <= max |x|, so the bound a GEMM / conv
    epilogue published for ``x`` also scales the result when it feeds the next fp16 GEMM (an upper bound is all the
    power-of-two operand scale needs) -- no streaming maximum pass over the activation plane."""
    out = F.leaky_relu(x, slope) if slope else F.relu(x)
    slot = _absmax.get(x) if x.is_cuda else None
    if slot is not None:
        publish_absmax(out, slot)
    return out


def leaky_relu(x):
    return relu(x, 0.01)  # F.leaky_relu's default slope


def _rowmajor(t):
    """2-D fp32 CUDA view usable by TMA: unit column stride, 16-byte aligned base and pitch."""
    if t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def _launch_fwd(x1, x2, split, n_out, bias, relu_in, mask, residual, out):
    rows, k1 = x1.shape
    k2 = 0 if x2 is None else x2.shape[1]
    if len(split) == 3:  # fp16 flavour: operand maximum first, then the GEMM
        w_hi, w_lo, w_slot = split
        x_slot = operand_absmax(x1, x2)
        out_slot = torch.empty(1, dtype=torch.int32, device=x1.device) if out.is_contiguous() else None
        _lib.call("t2h_linear_fwd_f16", ptr(x1), x1.stride(0), k1, ptr(x2), 0 if x2 is None else x2.stride(0), k2, rows,
                  ptr(x_slot), ptr(w_hi), ptr(w_lo), ptr(w_slot), n_out, ptr(bias), int(relu_in), ptr(mask),
                  0 if mask is None else mask.stride(0), ptr(residual), 0 if residual is None else residual.stride(0),
                  ptr(out), out.stride(0), ptr(out_slot))
        if out_slot is not None:
            publish_absmax(out, out_slot)
        return
    w_hi, w_lo = split
    _lib.call("t2h_linear_fwd", ptr(x1), x1.stride(0), k1, ptr(x2), 0 if x2 is None else x2.stride(0), k2, rows,
              ptr(w_hi), ptr(w_lo), n_out, ptr(bias), int(relu_in), ptr(mask), 0 if mask is None else mask.stride(0),
              ptr(residual), 0 if residual is None else residual.stride(0), ptr(out), out.stride(0))


def _launch_wgrad(gy, x, relu_in, grad_w_view, grad_b=None):
    rows, n_out = gy.shape
    k_in = x.shape[1]
    lib = _lib.load()
    ws_bytes = int(lib.t2h_linear_wgrad_workspace_bytes(rows, n_out, k_in))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=gy.device)
    if USE_F16_WGRAD and k_in >= 128 and n_out >= 32:
        _lib.call("t2h_linear_wgrad_f16", ptr(gy), gy.stride(0), ptr(operand_absmax(gy)), ptr(x), x.stride(0),
                  ptr(operand_absmax(x)), rows, n_out, k_in, int(relu_in), ptr(ws), ws_bytes, ptr(grad_w_view),
                  grad_w_view.stride(0), ptr(grad_b))
        return
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


# ------------------------------------------------------------------------------------------
# t2h::linear / t2h::linear_bwd -- nn.Linear (+ ReLU-on-load, concat, residual) and its backward as torch custom ops
# over t2h_linear_fwd[_f16] / t2h_linear_wgrad[_f16] / t2h_colsum (t2h_linear_fwd / t2h_linear_bwd)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::linear", mutates_args=())
def _linear_op(x1: Tensor, x2: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor],
               relu_in: bool) -> Tensor:
    """relu?([x1 | x2]) @ weight.T + bias + residual on 2-D row-major operands."""
    x1 = _rowmajor(x1)
    x2 = None if x2 is None else _rowmajor(x2)
    residual = None if residual is None else _rowmajor(residual)
    n_out = weight.shape[0]
    split = _cache.get(weight, f16=use_f16(n_out, weight.shape[1]))
    out = torch.empty(x1.shape[0], n_out, dtype=torch.float32, device=x1.device)
    _launch_fwd(x1, x2, split, n_out, bias, relu_in, None, residual, out)
    return out


@_linear_op.register_fake
def _(x1, x2, weight, bias, residual, relu_in):
    return x1.new_empty(x1.shape[0], weight.shape[0])


@torch.library.custom_op("t2h::linear_bwd", mutates_args=())
def _linear_bwd_op(gy: Tensor, x1: Tensor, x2: Optional[Tensor], weight: Tensor, relu_in: bool, has_bias: bool,
                   need_x1: bool, need_x2: bool, need_w: bool, need_b: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(d_x1, d_x2, d_weight, d_bias); a gradient that is not needed comes back as an empty tensor."""
    gy = _rowmajor(gy)
    x1 = _rowmajor(x1)
    x2 = None if x2 is None else _rowmajor(x2)
    k1 = x1.shape[1]
    n_out, k_total = weight.shape
    d_x1, d_x2, d_w, d_b = (gy.new_empty(0) for _ in range(4))  # distinct tensors: op outputs must not alias
    if need_x1:
        split = _cache.get(weight, 0, k1, transposed=True, f16=use_f16(k1, n_out))
        d_x1 = torch.empty_like(x1)
        _launch_fwd(gy, None, split, k1, None, False, x1 if relu_in else None, None, d_x1)
    if x2 is not None and need_x2:
        split = _cache.get(weight, k1, k_total, transposed=True, f16=use_f16(k_total - k1, n_out))
        d_x2 = torch.empty_like(x2)
        _launch_fwd(gy, None, split, k_total - k1, None, False, x2 if relu_in else None, None, d_x2)
    want_bias = has_bias and need_b
    if need_w:
        d_w = torch.empty(n_out, k_total, dtype=torch.float32, device=gy.device)
        if want_bias:  # the bias gradient (column sum of gy) rides along with the first wgrad launch
            d_b = torch.empty(n_out, dtype=torch.float32, device=gy.device)
        _launch_wgrad(gy, x1, relu_in, d_w[:, :k1], d_b if want_bias else None)
        if x2 is not None:
            _launch_wgrad(gy, x2, relu_in, d_w[:, k1:])
    elif want_bias:
        d_b = colsum(gy)
    return d_x1, d_x2, d_w, d_b


@_linear_bwd_op.register_fake
def _(gy, x1, x2, weight, relu_in, has_bias, need_x1, need_x2, need_w, need_b):
    e = gy.new_empty(0)
    return (torch.empty_like(x1) if need_x1 else e, torch.empty_like(x2) if (x2 is not None and need_x2) else e,
            torch.empty_like(weight) if need_w else e, gy.new_empty(weight.shape[0]) if (has_bias and need_b) else e)


def _linear_setup(ctx, inputs, output):
    x1, x2, weight, bias, residual, relu_in = inputs
    ctx.save_for_backward(x1, x2, weight)
    ctx.relu_in, ctx.has_bias = relu_in, bias is not None


def _linear_backward(ctx, gy):
    x1, x2, weight = ctx.saved_tensors
    need = ctx.needs_input_grad
    d_x1, d_x2, d_w, d_b = torch.ops.t2h.linear_bwd(gy, x1, x2, weight, ctx.relu_in, ctx.has_bias, need[0],
                                                    x2 is not None and need[1], need[2], ctx.has_bias and need[3])
    pick = lambda t: t if t.numel() else None
    return pick(d_x1), pick(d_x2), pick(d_w), pick(d_b), (gy if need[4] else None), None


_linear_op.register_autograd(_linear_backward, setup_context=_linear_setup)


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
    y = torch.ops.t2h.linear(x2d, x22d, weight, bias, res2d, bool(relu_in))
    y = y.view(*lead, weight.shape[0])
    return y[..., :n_out] if pn else y
