"""Plane convolutions on the tcgen05 3xTF32 pipeline.

The reference's plane CNN (alto.py:59-61,74,98-114,157-182,215-236,354,380) and ConvDecoder
(pixel.py:17-32) are ``nn.Conv2d`` / ``nn.ConvTranspose2d`` modules on cuDNN.  With strict fp32 they run
on SIMT kernels (~25 TFLOP/s on B200) and became the largest term of the step once the point path was
on tensor cores, so the three shapes they use are mapped onto the GEMM kernels of ``t2h_linear.cu``:

* 3x3 / padding 1  -> implicit GEMM, K chunk = (tap, 32-channel slice) fetched by 4-D TMA from the
                      tap-shifted pixel patch (zero fill outside the plane = the padding)
* 1x1              -> a linear layer over the pixels of the channels-last plane
* transposed 2x2 / stride 2 -> a linear layer to 4*Cout followed by a pixel shuffle

Parameters stay the modules' own fp32 ``weight`` / ``bias`` (checkpoints unchanged).  Tensors are the
logical (B, C, H, W) views with channels_last strides used everywhere in the model.  Shapes the
kernels do not cover (Cin or Cout not a multiple of 32, planes narrower than 16) fall back to cuDNN.
"""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib
from ._lib import ptr
from . import linear as _linear
from .linear import _cache, linear, operand_absmax, publish_absmax, USE_LIBRARY_GEMM


def _fwd_matrix(w):      # (Cout, Cin, 3, 3) -> [Cout, (ky, kx, ci)]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _dgrad_matrix(w):    # -> [Cin, (ky', kx', co)] with the taps mirrored
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1)


def _launch_conv(x, split, cout, bias, relu_in, mask, out):
    B, H, W, cin = x.shape
    if len(split) == 3:  # fp16 flavour (linear.py): operand maxima in, output maximum out
        w_hi, w_lo, w_slot = split
        x_slot = operand_absmax(x.view(-1, cin), owner=x)
        out_slot = torch.empty(1, dtype=torch.int32, device=x.device)
        _lib.call("t2h_conv3x3_fwd_f16", ptr(x), B, H, W, cin, ptr(x_slot), ptr(w_hi), ptr(w_lo), ptr(w_slot), cout, ptr(bias),
                  int(relu_in), ptr(mask), None, ptr(out), ptr(out_slot))
        publish_absmax(out, out_slot)
        return
    w_hi, w_lo = split
    _lib.call("t2h_conv3x3_fwd", ptr(x), B, H, W, cin, ptr(w_hi), ptr(w_lo), cout, ptr(bias), int(relu_in), ptr(mask),
              None, ptr(out))


# t2h::conv3x3 / t2h::conv3x3_bwd -- the 3x3 / padding 1 convolution and its backward as torch custom ops over
# t2h_conv3x3_fwd[_f16] / t2h_conv3x3_wgrad[_f16]
@torch.library.custom_op("t2h::conv3x3", mutates_args=())
def _conv3x3_op(x: Tensor, weight: Tensor, bias: Optional[Tensor], relu_in: bool) -> Tensor:
    """x: channels-last (B, H, W, Cin) contiguous; returns (B, H, W, Cout)."""
    x = x.contiguous()
    B, H, W, cin = x.shape
    cout = weight.shape[0]
    split = _cache.get_matrix(weight, "conv3x3_fwd", _fwd_matrix, f16=_linear.USE_F16)
    out = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
    _launch_conv(x, split, cout, bias, relu_in, None, out)
    return out


@_conv3x3_op.register_fake
def _(x, weight, bias, relu_in):
    return x.new_empty(x.shape[0], x.shape[1], x.shape[2], weight.shape[0])


@torch.library.custom_op("t2h::conv3x3_bwd", mutates_args=())
def _conv3x3_bwd_op(g: Tensor, x: Tensor, weight: Tensor, relu_in: bool, has_bias: bool, need_x: bool,
                    need_w: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(d_x, d_weight (Cout, Cin, 3, 3), d_bias); a gradient that is not needed comes back empty."""
    g = g.contiguous()
    x = x.contiguous()
    B, H, W, cin = x.shape
    cout = weight.shape[0]
    d_x, d_w, d_b = (g.new_empty(0) for _ in range(3))  # distinct tensors: op outputs must not alias
    if need_x:
        split = _cache.get_matrix(weight, "conv3x3_dgrad", _dgrad_matrix, f16=_linear.USE_F16)
        d_x = torch.empty_like(x)
        _launch_conv(g, split, cin, None, False, x if relu_in else None, d_x)
    if need_w or has_bias:
        lib = _lib.load()
        ws_bytes = int(lib.t2h_conv3x3_wgrad_workspace_bytes(B, H, W, cin, cout))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
        d_wm = torch.empty(cout, 9 * cin, dtype=torch.float32, device=g.device)
        if has_bias:
            d_b = torch.empty(cout, dtype=torch.float32, device=g.device)
        if _linear.USE_F16_WGRAD:
            _lib.call("t2h_conv3x3_wgrad_f16", ptr(g), ptr(operand_absmax(g.view(-1, cout), owner=g)), ptr(x),
                      ptr(operand_absmax(x.view(-1, cin), owner=x)), B, H, W, cin, cout, int(relu_in), ptr(ws),
                      ws_bytes, ptr(d_wm), ptr(d_b) if has_bias else None)
        else:
            _lib.call("t2h_conv3x3_wgrad", ptr(g), ptr(x), B, H, W, cin, cout, int(relu_in), ptr(ws), ws_bytes,
                      ptr(d_wm), ptr(d_b) if has_bias else None)
        d_w = d_wm.view(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    return d_x, d_w, d_b


@_conv3x3_bwd_op.register_fake
def _(g, x, weight, relu_in, has_bias, need_x, need_w):
    e = g.new_empty(0)
    return (torch.empty_like(x) if need_x else e, torch.empty_like(weight) if (need_w or has_bias) else e,
            g.new_empty(weight.shape[0]) if has_bias else e)


def _conv3x3_setup(ctx, inputs, output):
    x, weight, bias, relu_in = inputs
    ctx.save_for_backward(x, weight)
    ctx.relu_in, ctx.has_bias = relu_in, bias is not None


def _conv3x3_backward(ctx, g):
    x, weight = ctx.saved_tensors
    need = ctx.needs_input_grad
    d_x, d_w, d_b = torch.ops.t2h.conv3x3_bwd(g, x, weight, ctx.relu_in, ctx.has_bias and need[2], need[0], need[1])
    pick = lambda t: t if t.numel() else None
    return pick(d_x), (pick(d_w) if need[1] else None), pick(d_b), None


_conv3x3_op.register_autograd(_conv3x3_backward, setup_context=_conv3x3_setup)


def _tc_ok_3x3(x, weight):
    B, C, H, W = x.shape
    return (not USE_LIBRARY_GEMM and x.is_cuda and x.dtype == torch.float32 and weight.shape[2:] == (3, 3)
            and C % 32 == 0 and weight.shape[0] % 32 == 0 and W % 16 == 0 and H % 8 == 0)


def conv3x3(x, weight, bias=None, relu_in=False):
    """F.conv2d(relu?(x), weight, bias, padding=1) on logical (B, C, H, W) tensors."""
    if not _tc_ok_3x3(x, weight):
        return F.conv2d(F.relu(x) if relu_in else x, weight, bias, padding=1)
    y = torch.ops.t2h.conv3x3(x.permute(0, 2, 3, 1).contiguous(), weight, bias, bool(relu_in))
    return y.permute(0, 3, 1, 2)


def conv1x1(x, weight, bias=None, relu_in=False, residual=None):
    """F.conv2d(relu?(x), weight, bias) (+ residual) for 1x1 kernels = a linear layer over the pixels; the residual
    add (alto.py:111-114: ``x = x + conv1x1(x_after_conv)``) rides in the GEMM epilogue."""
    if USE_LIBRARY_GEMM or not x.is_cuda:
        y = F.conv2d(F.relu(x) if relu_in else x, weight, bias)
        return y if residual is None else y + residual
    y = linear(x.permute(0, 2, 3, 1), weight.reshape(weight.shape[0], weight.shape[1]), bias, relu_in=relu_in,
               residual=None if residual is None else residual.permute(0, 2, 3, 1))
    return y.permute(0, 3, 1, 2)


def conv_transpose2x2(x, weight, bias=None):
    """F.conv_transpose2d(x, weight, bias, stride=2) for 2x2 kernels: out[2y+dy, 2x+dx] = x[y, x] @ W[:, :, dy, dx]."""
    if USE_LIBRARY_GEMM or not x.is_cuda:
        return F.conv_transpose2d(x, weight, bias, stride=2)
    B, cin, H, W = x.shape
    cout = weight.shape[1]
    wm = weight.permute(2, 3, 1, 0).reshape(4 * cout, cin)           # rows ordered (dy, dx, co)
    bm = None if bias is None else bias.repeat(4)
    y = linear(x.permute(0, 2, 3, 1), wm, bm)                          # (B, H, W, 4*cout)
    y = y.view(B, H, W, 2, 2, cout).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, cout)
    return y.permute(0, 3, 1, 2)


def apply_conv(module, x, relu_in=False, residual=None):
    """Run an ``nn.Conv2d`` / ``nn.ConvTranspose2d`` module of the plane CNN through the kernels above
    (its parameters are used as they are); anything else is called as a module.  ``residual`` is added to the
    result -- inside the GEMM epilogue for 1x1 convolutions."""
    if isinstance(module, torch.nn.Conv2d) and module.groups == 1 and module.stride == (1, 1) and module.dilation == (1, 1):
        if module.kernel_size == (1, 1) and module.padding == (0, 0):
            return conv1x1(x, module.weight, module.bias, relu_in=relu_in, residual=residual)
        if module.kernel_size == (3, 3) and module.padding == (1, 1) and module.padding_mode == 'zeros':
            y = conv3x3(x, module.weight, module.bias, relu_in=relu_in)
            return y if residual is None else y + residual
    if (isinstance(module, torch.nn.ConvTranspose2d) and module.kernel_size == (2, 2) and module.stride == (2, 2)
            and module.padding == (0, 0) and module.output_padding == (0, 0) and module.groups == 1 and not relu_in):
        y = conv_transpose2x2(x, module.weight, module.bias)
        return y if residual is None else y + residual
    y = module(F.relu(x) if relu_in else x)
    return y if residual is None else y + residual
