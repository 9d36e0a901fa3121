"""Autograd wrappers around the C ABI (include/t2h.h).

Each Function's forward / backward is one kernel launch through ctypes; tensors are allocated by
PyTorch, the C side only sees raw pointers and the current stream.  Backward runs on the autograd
thread -- the C entry points are re-entrant and stateless.

Layouts: per-point features are (n_rows, C) row-major; planes are channels-last (B, r, r, C).
"""
import torch

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


def _geom(level):
    return (ptr(level.perm), ptr(level.cell_start), level.n_seg, level.shift)


class _SegMaxPool(torch.autograd.Function):
    """pool_local with scatter_max (pointnet.py:92-99): per-cell max broadcast back to the points."""

    @staticmethod
    def forward(ctx, rows, level, want_plane):
        rows = _rows(rows, "seg_max_pool")
        C = rows.shape[1]
        pooled = torch.empty_like(rows)
        arg = torch.empty(level.n_seg, C, dtype=torch.int32, device=rows.device)
        plane = torch.empty(level.n_seg, C, dtype=torch.float32, device=rows.device) if want_plane else None
        call("t2h_seg_max_fwd", ptr(rows), ptr(level.perm), ptr(level.tie), *_geom(level)[1:], C, level.morton, level.reso, ptr(pooled), ptr(plane), ptr(arg))
        ctx.level = level
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        if want_plane:
            ctx.mark_non_differentiable(plane)
            return pooled, arg, plane
        return pooled, arg

    @staticmethod
    def backward(ctx, g_pooled, *_unused):
        (arg,) = ctx.saved_tensors
        level = ctx.level
        g_pooled = g_pooled.contiguous()
        C = g_pooled.shape[1]
        g_rows = torch.empty_like(g_pooled)
        call("t2h_seg_max_bwd", ptr(g_pooled), None, *_geom(level), C, level.morton, level.reso, ptr(arg), ptr(g_rows))
        return g_rows, None, None


class _SegMaxPlane(torch.autograd.Function):
    """torch_scatter.scatter_max proper: per-cell max + argmax (no gather-back)."""

    @staticmethod
    def forward(ctx, rows, level):
        rows = _rows(rows, "seg_max_plane")
        C = rows.shape[1]
        arg = torch.empty(level.n_seg, C, dtype=torch.int32, device=rows.device)
        plane = torch.empty(level.n_seg, C, dtype=torch.float32, device=rows.device)
        call("t2h_seg_max_fwd", ptr(rows), ptr(level.perm), ptr(level.tie), *_geom(level)[1:], C, level.morton, level.reso, None, ptr(plane), ptr(arg))
        ctx.level = level
        ctx.n_rows = rows.shape[0]
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return plane, arg

    @staticmethod
    def backward(ctx, g_plane, _g_arg):
        (arg,) = ctx.saved_tensors
        level = ctx.level
        g_plane = g_plane.contiguous()
        C = g_plane.shape[1]
        g_rows = torch.empty(ctx.n_rows, C, dtype=torch.float32, device=g_plane.device)
        call("t2h_seg_max_bwd", None, ptr(g_plane), *_geom(level), C, level.morton, level.reso, ptr(arg), ptr(g_rows))
        return g_rows, None


class _SegReduce(torch.autograd.Function):
    """scatter_mean / scatter_sum onto the plane (pointnet.py:101-111, alto.py:76-88,187-197)."""

    @staticmethod
    def forward(ctx, rows, level, mean):
        rows = _rows(rows, "seg_reduce")
        C = rows.shape[1]
        plane = torch.empty(level.n_seg, C, dtype=torch.float32, device=rows.device)
        call("t2h_seg_reduce_fwd", ptr(rows), rows.shape[0], *_geom(level), C, level.morton, level.reso, int(mean), ptr(plane))
        ctx.level, ctx.mean, ctx.n_rows = level, mean, rows.shape[0]
        return plane

    @staticmethod
    def backward(ctx, g_plane):
        level = ctx.level
        g_plane = g_plane.contiguous()
        C = g_plane.shape[1]
        g_rows = torch.empty(ctx.n_rows, C, dtype=torch.float32, device=g_plane.device)
        call("t2h_seg_broadcast", ptr(g_plane), *_geom(level), C, level.morton, level.reso, int(ctx.mean), ptr(g_rows))
        return g_rows, None, None


class _SegBroadcast(torch.autograd.Function):
    """rows[i] = plane[cell(i)] (/count): the gather-back of pool_local for scatter_type='mean'."""

    @staticmethod
    def forward(ctx, plane, level, mean, n_rows):
        plane = _rows(plane, "seg_broadcast")
        C = plane.shape[1]
        rows = torch.empty(n_rows, C, dtype=torch.float32, device=plane.device)
        call("t2h_seg_broadcast", ptr(plane), *_geom(level), C, level.morton, level.reso, int(mean), ptr(rows))
        ctx.level, ctx.mean = level, mean
        return rows

    @staticmethod
    def backward(ctx, g_rows):
        level = ctx.level
        g_rows = g_rows.contiguous()
        C = g_rows.shape[1]
        g_plane = torch.empty(level.n_seg, C, dtype=torch.float32, device=g_rows.device)
        call("t2h_seg_reduce_fwd", ptr(g_rows), g_rows.shape[0], *_geom(level), C, level.morton, level.reso, int(ctx.mean), ptr(g_plane))
        return g_plane, None, None, None


class _BilinearSample(torch.autograd.Function):
    """F.grid_sample(plane, 2p-1, bilinear, border, align_corners=True) (alto.py:90-95,199-205)."""

    @staticmethod
    def forward(ctx, plane, level):
        _lib.require_cuda_f32(plane, "bilinear_sample(plane)")
        if plane.dim() != 4 or plane.shape[1] != level.reso or plane.shape[2] != level.reso:
            raise RuntimeError(f"bilinear_sample: expected channels-last (B, {level.reso}, {level.reso}, C), got {tuple(plane.shape)}")
        if plane.shape[3] not in SUPPORTED_C:
            raise RuntimeError(f"bilinear_sample: unsupported channel count {plane.shape[3]}")
        plane = plane.contiguous()
        C = plane.shape[3]
        n = level.n_points
        out = torch.empty(n, C, dtype=torch.float32, device=plane.device)
        xyz = level.xyz_sorted
        call("t2h_bilinear_sample_fwd", ptr(plane), level.reso, C, ptr(xyz), xyz.shape[1], ptr(level.perm),
             ptr(level.tile_ids), n, level.N or 1, ptr(out))
        ctx.level = level
        ctx.shape = tuple(plane.shape)
        return out

    @staticmethod
    def backward(ctx, g_rows):
        level = ctx.level
        g_rows = g_rows.contiguous()
        C = g_rows.shape[1]
        g_plane = torch.empty(ctx.shape, dtype=torch.float32, device=g_rows.device)
        xyz = level.xyz_sorted
        ws_bytes = int(_lib.load().t2h_bilinear_sample_bwd_workspace_bytes(level.reso, C, level.n_seg, level.morton))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g_rows.device)
        call("t2h_bilinear_sample_bwd", ptr(g_rows), g_rows.shape[0], level.reso, C, ptr(xyz), xyz.shape[1], ptr(level.perm),
             ptr(level.cell_start), level.n_seg, level.shift, level.morton, ptr(ws), ws_bytes, ptr(g_plane))
        return g_plane, None


class _UpsampleBilinear(torch.autograd.Function):
    """F.interpolate(plane, size, bilinear, align_corners=True) on channels-last planes (pixel.py:105-111)."""

    @staticmethod
    def forward(ctx, plane, out_h, out_w):
        _lib.require_cuda_f32(plane, "upsample_bilinear(plane)")
        if plane.dim() != 4 or plane.shape[3] not in SUPPORTED_C:
            raise RuntimeError(f"upsample_bilinear: expected channels-last (B, h, w, C), got {tuple(plane.shape)}")
        plane = plane.contiguous()
        B, h, w, C = plane.shape
        out = torch.empty(B, out_h, out_w, C, dtype=torch.float32, device=plane.device)
        call("t2h_upsample_bilinear_fwd", ptr(plane), B, h, w, C, out_h, out_w, ptr(out))
        ctx.dims = (B, h, w, C, out_h, out_w)
        return out

    @staticmethod
    def backward(ctx, g_out):
        B, h, w, C, out_h, out_w = ctx.dims
        g_out = g_out.contiguous()
        g_in = torch.empty(B, h, w, C, dtype=torch.float32, device=g_out.device)
        call("t2h_upsample_bilinear_bwd", ptr(g_out), B, h, w, C, out_h, out_w, ptr(g_in))
        return g_in, None, None


# ------------------------------------------------------------------------------------------
# public functional API
# ------------------------------------------------------------------------------------------
def seg_max_pool(rows, level, return_arg=False):
    """Per-cell max of (n_rows, C) features, broadcast back to every point of the cell."""
    pooled, arg = _SegMaxPool.apply(rows, level, False)
    return (pooled, arg) if return_arg else pooled


def seg_max(rows, level):
    """(plane (n_seg, C), arg (n_seg, C) int32 row index or -1)"""
    return _SegMaxPlane.apply(rows, level)


def seg_mean(rows, level):
    """Per-cell mean -> (n_seg, C); empty cells are 0."""
    return _SegReduce.apply(rows, level, True)


def seg_sum(rows, level):
    return _SegReduce.apply(rows, level, False)


def seg_broadcast(plane, level, mean=False):
    """rows[i] = plane[cell(i)] (divided by the cell count when mean=True)."""
    return _SegBroadcast.apply(plane, level, mean, level.n_points)


def bilinear_sample(plane_cl, level):
    """plane_cl (B, r, r, C) channels-last -> (B*N, C) rows."""
    return _BilinearSample.apply(plane_cl, level)


def upsample_bilinear(plane_cl, size):
    """(B, h, w, C) -> (B, size, size, C), align_corners=True."""
    if isinstance(size, int):
        size = (size, size)
    return _UpsampleBilinear.apply(plane_cl, int(size[0]), int(size[1]))


def cell_index(xy: torch.Tensor, reso: int) -> torch.Tensor:
    """coordinate2index (utils/coordinate.py:12-28): (B, N, 2) -> (B, 1, N) int64."""
    _lib.require_cuda_f32(xy, "cell_index")
    if xy.dim() != 3 or xy.shape[2] < 2:
        raise RuntimeError(f"cell_index: expected (B, N, 2), got {tuple(xy.shape)}")
    xy = xy.contiguous()
    B, N, D = xy.shape
    out = torch.empty(B, 1, N, dtype=torch.int64, device=xy.device)
    call("t2h_cell_index", ptr(xy), B * N, D, int(reso), ptr(out))
    return out


def plane_to_nchw(plane_rows, B, reso):
    """(B*r*r, C) channels-last rows -> logical (B, C, r, r) view with channels_last strides."""
    C = plane_rows.shape[1]
    return plane_rows.view(B, reso, reso, C).permute(0, 3, 1, 2)


def nchw_to_plane(x):
    """logical (B, C, h, w) -> channels-last (B, h, w, C) contiguous (free if already channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()
