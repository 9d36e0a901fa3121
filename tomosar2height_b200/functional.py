"""Torch custom-op layer over the C ABI (include/t2h.h) for the segment / sampling operators.

Every operator is a ``torch.library.custom_op`` in the ``t2h::`` namespace with a ``register_fake``
shape function and a ``register_autograd`` formula whose backward is again a ``t2h::`` op; the
implementation of each op is ONE ctypes call into libt2h.so (tensors are allocated by PyTorch, the C side
only sees raw pointers, sizes and the current stream).  Backward runs on the autograd thread -- the C
entry points are re-entrant and stateless.  There is no CPU implementation: a CPU tensor raises.

Layouts: per-point features are (n_rows, C) row-major; planes are channels-last (B, r, r, C).
A "level" (topology.CellLevel / IndexLevel) carries the sort of the points: ``perm`` (sorted position ->
row, None when rows are stored sorted), ``keys`` (sort key of every sorted position), ``cell_start``.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import ptr, call

SUPPORTED_C = (4, 8, 16, 32, 64, 128, 256, 384, 512, 640, 768, 896, 1024)


def _rows(t, what):
    _lib.require_cuda_f32(t, what)
    if t.dim() != 2:
        raise RuntimeError(f"{what}: expected (n_rows, C), got {tuple(t.shape)}")
    if t.shape[1] not in SUPPORTED_C:
        raise RuntimeError(f"{what}: unsupported channel count {t.shape[1]} (supported: {SUPPORTED_C})")
    return t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def _seg_ws(n_rows, n_seg, C, device):
    ws = _workspace(_lib.load().t2h_seg_workspace_bytes(n_rows, n_seg, C), device)
    return ws, ws.numel()


# ------------------------------------------------------------------------------------------
# t2h::seg_max  -- torch_scatter.scatter_max (+ gather-back): pointnet.py:92-99
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::seg_max", mutates_args=())
def _seg_max(rows: Tensor, perm: Optional[Tensor], tie: Optional[Tensor], keys: Tensor, cell_start: Tensor,
             n_seg: int, shift: int, morton: int, reso: int, want_pooled: bool) -> Tuple[Tensor, Tensor, Tensor]:
    rows = _rows(rows, "seg_max")
    n, C = rows.shape
    plane = torch.empty(n_seg, C, dtype=torch.float32, device=rows.device)
    arg = torch.empty(n_seg, C, dtype=torch.int32, device=rows.device)
    pooled = torch.empty_like(rows) if want_pooled else rows.new_empty(0, C)
    ws, nb = _seg_ws(n, n_seg, C, rows.device)
    call("t2h_seg_max_fwd", ptr(rows), n, ptr(perm), ptr(tie), ptr(keys), ptr(cell_start), n_seg, shift, C, morton, reso,
         ptr(ws), nb, ptr(pooled) if want_pooled else None, ptr(plane), ptr(arg))
    return pooled, plane, arg


@_seg_max.register_fake
def _(rows, perm, tie, keys, cell_start, n_seg, shift, morton, reso, want_pooled):
    C = rows.shape[1]
    return (rows.new_empty(rows.shape if want_pooled else (0, C)), rows.new_empty(n_seg, C),
            rows.new_empty(n_seg, C, dtype=torch.int32))


@torch.library.custom_op("t2h::seg_max_bwd", mutates_args=())
def _seg_max_bwd(grad_pooled: Optional[Tensor], grad_plane: Optional[Tensor], arg: Tensor, n_rows: int,
                 perm: Optional[Tensor], keys: Tensor, cell_start: Tensor, n_seg: int, shift: int, morton: int,
                 reso: int) -> Tensor:
    C = arg.shape[1]
    gp = None if grad_pooled is None else _rows(grad_pooled, "seg_max_bwd(grad_pooled)")
    gl = None if grad_plane is None else _rows(grad_plane, "seg_max_bwd(grad_plane)")
    out = torch.empty(n_rows, C, dtype=torch.float32, device=arg.device)
    if gp is None and gl is None:
        return out.zero_()
    ws, nb = _seg_ws(n_rows, n_seg, C, arg.device) if gp is not None else (None, 0)
    call("t2h_seg_max_bwd", ptr(gp), ptr(gl), n_rows, ptr(perm), ptr(keys), ptr(cell_start), n_seg, shift, C, morton, reso,
         ptr(arg), ptr(ws), nb, ptr(out))
    return out


@_seg_max_bwd.register_fake
def _(grad_pooled, grad_plane, arg, n_rows, perm, keys, cell_start, n_seg, shift, morton, reso):
    return arg.new_empty(n_rows, arg.shape[1], dtype=torch.float32)


def _seg_max_setup(ctx, inputs, output):
    rows, perm, tie, keys, cell_start, n_seg, shift, morton, reso, want_pooled = inputs
    ctx.set_materialize_grads(False)  # an unused output (plane of pool_local) must not cost a zero plane
    ctx.save_for_backward(output[2], perm, keys, cell_start)
    ctx.meta = (rows.shape[0], n_seg, shift, morton, reso, want_pooled)


def _seg_max_backward(ctx, g_pooled, g_plane, _g_arg):
    arg, perm, keys, cell_start = ctx.saved_tensors
    n_rows, n_seg, shift, morton, reso, want_pooled = ctx.meta
    g = torch.ops.t2h.seg_max_bwd(g_pooled.contiguous() if (want_pooled and g_pooled is not None) else None,
                                  None if g_plane is None else g_plane.contiguous(), arg, n_rows, perm, keys, cell_start,
                                  n_seg, shift, morton, reso)
    return g, None, None, None, None, None, None, None, None, None


_seg_max.register_autograd(_seg_max_backward, setup_context=_seg_max_setup)


# ------------------------------------------------------------------------------------------
# t2h::seg_reduce / t2h::seg_broadcast -- torch_scatter.scatter_mean and its backward:
# pointnet.py:101-111, alto.py:76-88,187-197
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::seg_reduce", mutates_args=())
def _seg_reduce(rows: Tensor, perm: Optional[Tensor], keys: Tensor, cell_start: Tensor, n_seg: int, shift: int,
                morton: int, reso: int, mean: bool) -> Tensor:
    rows = _rows(rows, "seg_reduce")
    n, C = rows.shape
    plane = torch.empty(n_seg, C, dtype=torch.float32, device=rows.device)
    ws, nb = _seg_ws(n, n_seg, C, rows.device)
    call("t2h_seg_reduce_fwd", ptr(rows), n, ptr(perm), ptr(keys), ptr(cell_start), n_seg, shift, C, morton, reso, int(mean),
         ptr(ws), nb, ptr(plane))
    return plane


@_seg_reduce.register_fake
def _(rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean):
    return rows.new_empty(n_seg, rows.shape[1])


@torch.library.custom_op("t2h::seg_broadcast", mutates_args=())
def _seg_broadcast(plane: Tensor, n_rows: int, perm: Optional[Tensor], keys: Tensor, cell_start: Tensor, n_seg: int,
                   shift: int, morton: int, reso: int, mean: bool) -> Tensor:
    plane = _rows(plane, "seg_broadcast")
    C = plane.shape[1]
    rows = torch.empty(n_rows, C, dtype=torch.float32, device=plane.device)
    call("t2h_seg_broadcast", ptr(plane), n_rows, ptr(perm), ptr(keys), ptr(cell_start), n_seg, shift, C, morton, reso,
         int(mean), ptr(rows))
    return rows


@_seg_broadcast.register_fake
def _(plane, n_rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean):
    return plane.new_empty(n_rows, plane.shape[1])


def _seg_reduce_setup(ctx, inputs, output):
    rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean = inputs
    ctx.save_for_backward(perm, keys, cell_start)
    ctx.meta = (rows.shape[0], n_seg, shift, morton, reso, mean)


def _seg_reduce_backward(ctx, g_plane):
    perm, keys, cell_start = ctx.saved_tensors
    n_rows, n_seg, shift, morton, reso, mean = ctx.meta
    g = torch.ops.t2h.seg_broadcast(g_plane.contiguous(), n_rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean)
    return g, None, None, None, None, None, None, None, None


_seg_reduce.register_autograd(_seg_reduce_backward, setup_context=_seg_reduce_setup)


def _seg_broadcast_setup(ctx, inputs, output):
    plane, n_rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean = inputs
    ctx.save_for_backward(perm, keys, cell_start)
    ctx.meta = (n_seg, shift, morton, reso, mean)


def _seg_broadcast_backward(ctx, g_rows):
    perm, keys, cell_start = ctx.saved_tensors
    n_seg, shift, morton, reso, mean = ctx.meta
    g = torch.ops.t2h.seg_reduce(g_rows.contiguous(), perm, keys, cell_start, n_seg, shift, morton, reso, mean)
    return g, None, None, None, None, None, None, None, None, None


_seg_broadcast.register_autograd(_seg_broadcast_backward, setup_context=_seg_broadcast_setup)


@torch.library.custom_op("t2h::seg_broadcast_add", mutates_args=())
def _seg_broadcast_add(plane: Tensor, add_rows: Tensor, perm: Optional[Tensor], keys: Tensor, cell_start: Tensor, n_seg: int,
                       shift: int, morton: int, reso: int, mean: bool) -> Tuple[Tensor, Tensor]:
    """(rows = plane[cell] / count + add_rows, device word with the bit pattern of max |rows|)"""
    plane = _rows(plane, "seg_broadcast_add(plane)")
    add_rows = _rows(add_rows, "seg_broadcast_add(rows)")
    n, C = add_rows.shape
    rows = torch.empty_like(add_rows)
    slot = torch.empty(1, dtype=torch.int32, device=plane.device)
    call("t2h_seg_broadcast_add", ptr(plane), ptr(add_rows), n, ptr(perm), ptr(keys), ptr(cell_start), n_seg, shift, C, morton,
         reso, int(mean), ptr(rows), ptr(slot))
    return rows, slot


@_seg_broadcast_add.register_fake
def _(plane, add_rows, perm, keys, cell_start, n_seg, shift, morton, reso, mean):
    return torch.empty_like(add_rows), plane.new_empty(1, dtype=torch.int32)


class _SegMeanCarry(torch.autograd.Function):
    """``rows -> (mean plane, rows)``: the mean-scatter of alto.py:130 together with the hand-over of the same
    per-point tensor to the next level's fc_c (alto.py:127).  Forward = t2h::seg_reduce, the second output is the
    input itself.  Backward: the two gradient branches meet here, so the gather of the plane gradient, their sum and
    the operand maximum of the GEMMs that consume the sum are ONE kernel (t2h::seg_broadcast_add) instead of a
    broadcast, autograd's add and a maximum pass."""

    @staticmethod
    def forward(ctx, rows, level):
        ctx.level = level
        ctx.n_rows = rows.shape[0]
        plane = torch.ops.t2h.seg_reduce(rows, level.perm, *_lv(level), True)
        return plane, rows.view_as(rows)

    @staticmethod
    def backward(ctx, g_plane, g_rows):
        level = ctx.level
        if g_plane is None:
            return g_rows, None
        if g_rows is None:
            return torch.ops.t2h.seg_broadcast(g_plane.contiguous(), ctx.n_rows, level.perm, *_lv(level), True), None
        out, slot = torch.ops.t2h.seg_broadcast_add(g_plane.contiguous(), g_rows.contiguous(), level.perm, *_lv(level), True)
        from .linear import publish_absmax
        publish_absmax(out, slot)
        return out, None


# ------------------------------------------------------------------------------------------
# t2h::bilinear_sample -- F.grid_sample(plane, 2p-1, bilinear, border, align_corners=True): alto.py:90-95,199-205
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::bilinear_sample", mutates_args=())
def _bilinear_sample(plane: Tensor, xyz: Tensor, perm: Optional[Tensor], tile_ids: Optional[Tensor], keys: Tensor,
                     cell_start: Tensor, n_seg: int, shift: int, morton: int, n_per_batch: int) -> Tensor:
    _lib.require_cuda_f32(plane, "bilinear_sample(plane)")
    plane = plane.contiguous()
    B, reso, _, C = plane.shape
    n = xyz.shape[0]
    out = torch.empty(n, C, dtype=torch.float32, device=plane.device)
    call("t2h_bilinear_sample_fwd", ptr(plane), reso, C, ptr(xyz), xyz.shape[1], ptr(perm), ptr(tile_ids), n, n_per_batch,
         ptr(out))
    return out


@_bilinear_sample.register_fake
def _(plane, xyz, perm, tile_ids, keys, cell_start, n_seg, shift, morton, n_per_batch):
    return plane.new_empty(xyz.shape[0], plane.shape[3])


@torch.library.custom_op("t2h::bilinear_sample_bwd", mutates_args=())
def _bilinear_sample_bwd(grad_rows: Tensor, B: int, reso: int, xyz: Tensor, perm: Optional[Tensor], keys: Tensor,
                         cell_start: Tensor, n_seg: int, shift: int, morton: int) -> Tensor:
    g = _rows(grad_rows, "bilinear_sample_bwd")
    n, C = g.shape
    g_plane = torch.empty(B, reso, reso, C, dtype=torch.float32, device=g.device)
    nb = int(_lib.load().t2h_bilinear_sample_bwd_workspace_bytes(reso, C, n, n_seg, morton))
    ws = _workspace(nb, g.device)
    call("t2h_bilinear_sample_bwd", ptr(g), n, reso, C, ptr(xyz), xyz.shape[1], ptr(perm), ptr(keys), ptr(cell_start), n_seg,
         shift, morton, ptr(ws), nb, ptr(g_plane))
    return g_plane


@_bilinear_sample_bwd.register_fake
def _(grad_rows, B, reso, xyz, perm, keys, cell_start, n_seg, shift, morton):
    return grad_rows.new_empty(B, reso, reso, grad_rows.shape[1])


def _bilinear_sample_setup(ctx, inputs, output):
    plane, xyz, perm, tile_ids, keys, cell_start, n_seg, shift, morton, n_per_batch = inputs
    ctx.save_for_backward(xyz, perm, keys, cell_start)
    ctx.meta = (plane.shape[0], plane.shape[1], n_seg, shift, morton)


def _bilinear_sample_backward(ctx, g_rows):
    xyz, perm, keys, cell_start = ctx.saved_tensors
    B, reso, n_seg, shift, morton = ctx.meta
    g = torch.ops.t2h.bilinear_sample_bwd(g_rows.contiguous(), B, reso, xyz, perm, keys, cell_start, n_seg, shift, morton)
    return g, None, None, None, None, None, None, None, None, None


_bilinear_sample.register_autograd(_bilinear_sample_backward, setup_context=_bilinear_sample_setup)


# ------------------------------------------------------------------------------------------
# t2h::upsample_bilinear -- F.interpolate(plane, size, bilinear, align_corners=True): pixel.py:105-111
# (the §8(b) name "t2h_upsample2x" is this op at out = 2 * in; any size is supported)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::upsample_bilinear", mutates_args=())
def _upsample(plane: Tensor, out_h: int, out_w: int) -> Tensor:
    _lib.require_cuda_f32(plane, "upsample_bilinear(plane)")
    plane = plane.contiguous()
    B, h, w, C = plane.shape
    out = torch.empty(B, out_h, out_w, C, dtype=torch.float32, device=plane.device)
    call("t2h_upsample_bilinear_fwd", ptr(plane), B, h, w, C, out_h, out_w, ptr(out))
    return out


@_upsample.register_fake
def _(plane, out_h, out_w):
    return plane.new_empty(plane.shape[0], out_h, out_w, plane.shape[3])


@torch.library.custom_op("t2h::upsample_bilinear_bwd", mutates_args=())
def _upsample_bwd(grad_out: Tensor, h: int, w: int) -> Tensor:
    _lib.require_cuda_f32(grad_out, "upsample_bilinear_bwd")
    g = grad_out.contiguous()
    B, out_h, out_w, C = g.shape
    g_in = torch.empty(B, h, w, C, dtype=torch.float32, device=g.device)
    call("t2h_upsample_bilinear_bwd", ptr(g), B, h, w, C, out_h, out_w, ptr(g_in))
    return g_in


@_upsample_bwd.register_fake
def _(grad_out, h, w):
    return grad_out.new_empty(grad_out.shape[0], h, w, grad_out.shape[3])


def _upsample_setup(ctx, inputs, output):
    ctx.hw = (inputs[0].shape[1], inputs[0].shape[2])


def _upsample_backward(ctx, g_out):
    return torch.ops.t2h.upsample_bilinear_bwd(g_out.contiguous(), ctx.hw[0], ctx.hw[1]), None, None


_upsample.register_autograd(_upsample_backward, setup_context=_upsample_setup)


# ------------------------------------------------------------------------------------------
# t2h::cell_index -- coordinate2index (utils/coordinate.py:12-28)
# ------------------------------------------------------------------------------------------
@torch.library.custom_op("t2h::cell_index", mutates_args=())
def _cell_index(xy: Tensor, reso: int) -> Tensor:
    _lib.require_cuda_f32(xy, "cell_index")
    xy = xy.contiguous()
    B, N, D = xy.shape
    out = torch.empty(B, 1, N, dtype=torch.int64, device=xy.device)
    call("t2h_cell_index", ptr(xy), B * N, D, int(reso), ptr(out))
    return out


@_cell_index.register_fake
def _(xy, reso):
    return xy.new_empty(xy.shape[0], 1, xy.shape[1], dtype=torch.int64)


# ------------------------------------------------------------------------------------------
# public functional API (level objects -> op arguments)
# ------------------------------------------------------------------------------------------
def _lv(level):
    return level.keys, level.cell_start, level.n_seg, level.shift, level.morton, level.reso


def seg_max_pool(rows, level, return_arg=False):
    """Per-cell max of (n_rows, C) features, broadcast back to every point of the cell."""
    pooled, _plane, arg = torch.ops.t2h.seg_max(rows, level.perm, level.tie, *_lv(level), True)
    return (pooled, arg) if return_arg else pooled


def seg_max(rows, level):
    """(plane (n_seg, C), arg (n_seg, C) int32 row index or -1)"""
    _pooled, plane, arg = torch.ops.t2h.seg_max(rows, level.perm, level.tie, *_lv(level), False)
    return plane, arg


def seg_mean(rows, level):
    """Per-cell mean -> (n_seg, C); empty cells are 0."""
    return torch.ops.t2h.seg_reduce(rows, level.perm, *_lv(level), True)


def seg_mean_carry(rows, level):
    """(per-cell mean (n_seg, C), rows) for a tensor that feeds the mean-scatter AND a later consumer; see _SegMeanCarry."""
    if not (torch.is_grad_enabled() and rows.requires_grad):
        return seg_mean(rows, level), rows
    return _SegMeanCarry.apply(rows, level)


def seg_sum(rows, level):
    return torch.ops.t2h.seg_reduce(rows, level.perm, *_lv(level), False)


def seg_broadcast(plane, level, mean=False):
    """rows[i] = plane[cell(i)] (divided by the cell count when mean=True)."""
    return torch.ops.t2h.seg_broadcast(plane, level.n_points, level.perm, *_lv(level), bool(mean))


def bilinear_sample(plane_cl, level):
    """plane_cl (B, r, r, C) channels-last -> (B*N, C) rows."""
    if plane_cl.dim() != 4 or plane_cl.shape[1] != level.reso or plane_cl.shape[2] != level.reso:
        raise RuntimeError(f"bilinear_sample: expected channels-last (B, {level.reso}, {level.reso}, C), got {tuple(plane_cl.shape)}")
    if plane_cl.shape[3] not in SUPPORTED_C:
        raise RuntimeError(f"bilinear_sample: unsupported channel count {plane_cl.shape[3]}")
    return torch.ops.t2h.bilinear_sample(plane_cl, level.xyz_sorted, level.perm, level.tile_ids, level.keys, level.cell_start,
                                         level.n_seg, level.shift, level.morton, level.N or 1)


def upsample_bilinear(plane_cl, size):
    """(B, h, w, C) -> (B, size, size, C), align_corners=True."""
    if isinstance(size, int):
        size = (size, size)
    if plane_cl.dim() != 4 or plane_cl.shape[3] not in SUPPORTED_C:
        raise RuntimeError(f"upsample_bilinear: expected channels-last (B, h, w, C), got {tuple(plane_cl.shape)}")
    return torch.ops.t2h.upsample_bilinear(plane_cl, int(size[0]), int(size[1]))


def cell_index(xy: torch.Tensor, reso: int) -> torch.Tensor:
    """coordinate2index (utils/coordinate.py:12-28): (B, N, 2) -> (B, 1, N) int64."""
    if not isinstance(xy, torch.Tensor) or xy.dim() != 3 or xy.shape[2] < 2:
        raise RuntimeError(f"cell_index: expected (B, N, 2), got {tuple(xy.shape)}")
    return torch.ops.t2h.cell_index(xy, int(reso))


def plane_to_nchw(plane_rows, B, reso):
    """(B*r*r, C) channels-last rows -> logical (B, C, r, r) view with channels_last strides."""
    C = plane_rows.shape[1]
    return plane_rows.view(B, reso, reso, C).permute(0, 3, 1, 2)


def nchw_to_plane(x):
    """logical (B, C, h, w) -> channels-last (B, h, w, C) contiguous (free if already channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()
