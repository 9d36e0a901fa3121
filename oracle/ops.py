"""Operator-level CPU restatements (test infrastructure; see oracle/__init__.py).

Each function cites the reference call site it stands in for.  All functions are
plain PyTorch on CPU, work in fp32 or fp64, and are differentiable so that
``torch.autograd`` provides the backward the reference gets from autograd.
"""
import torch


def cell_index(xy: torch.Tensor, reso: int) -> torch.Tensor:
    """Cell id of every point; follows utils/coordinate.py:24-27.

    ``xy`` (B, N, 2) floating, normalised to the open interval (0, 1).
    Returns (B, 1, N) int64 with ``ix + reso * iy`` where ``i* = trunc(* x reso)``.
    The multiply happens in the dtype of ``xy`` (fp32 in the reference), the
    conversion truncates toward zero; there is no clamp.
    """
    cells = (xy * reso).long()
    flat = cells[:, :, 0] + reso * cells[:, :, 1]
    return flat[:, None, :]


def segment_max(src: torch.Tensor, index: torch.Tensor, dim_size: int):
    """torch_scatter.scatter_max(src, index, dim=-1, dim_size=M) on CPU.

    Stands in for pointnet.py:95.  ``src`` (B, C, N), ``index`` (B, 1, N) int64
    broadcast over C.  Rules (torch_scatter 2.1.x CPU): strict ``>``
    update => ties go to the smallest point index; an empty segment yields
    ``out = 0`` and ``arg = N``.  Backward routes the gradient to ``arg`` only.
    """
    B, C, N = src.shape
    idx = index.expand(B, C, N)
    with torch.no_grad():
        neg = torch.full((B, C, dim_size), float("-inf"), dtype=src.dtype)
        seg_max = neg.scatter_reduce(2, idx, src, "amax", include_self=True)
        hit = src == seg_max.gather(2, idx)
        pos = torch.arange(N, dtype=torch.int64).expand(B, C, N)
        cand = torch.where(hit, pos, torch.full_like(pos, N))
        arg = torch.full((B, C, dim_size), N, dtype=torch.int64)
        arg = arg.scatter_reduce(2, idx, cand, "amin", include_self=True)
        empty = arg == N
    picked = src.gather(2, arg.clamp(max=N - 1))
    out = torch.where(empty, torch.zeros((), dtype=src.dtype), picked)
    return out, arg


def segment_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter_mean(src, index, out=zeros(B, C, M)) on CPU.

    Stands in for pointnet.py:109, alto.py:85, alto.py:194.  Sum per cell, then
    divide by ``max(count, 1)``; empty cells stay 0.
    """
    B, C, N = src.shape
    idx = index.expand(B, C, N)
    total = torch.zeros((B, C, dim_size), dtype=src.dtype).scatter_add(2, idx, src)
    ones = torch.ones((B, 1, N), dtype=src.dtype)
    count = torch.zeros((B, 1, dim_size), dtype=src.dtype).scatter_add(2, index, ones)
    return total / count.clamp(min=1)


def bilinear_sample_points(plane: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    """F.grid_sample(plane, 2*xy-1, bilinear, padding_mode='border', align_corners=True).

    Stands in for alto.py:90-95 / alto.py:199-205.  ``plane`` (B, C, H, W),
    ``xy`` (B, N, 2) in (0, 1) with x -> W and y -> H.  Returns (B, C, N).
    Explicit four-tap arithmetic (ATen GridSampler semantics): unnormalise
    ``((g + 1) / 2) * (size - 1)``, clip to ``[0, size-1]``, corner weights from
    the distances to the opposite corner, taps outside the plane dropped.
    """
    B, C, H, W = plane.shape
    g = 2.0 * xy - 1.0
    ix = ((g[..., 0] + 1.0) / 2.0) * (W - 1)
    iy = ((g[..., 1] + 1.0) / 2.0) * (H - 1)
    ix = ix.clamp(0, W - 1)
    iy = iy.clamp(0, H - 1)
    x0 = ix.floor()
    y0 = iy.floor()
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = plane.reshape(B, C, H * W)

    def tap(xc, yc, w):
        ok = (xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)
        lin = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).long()
        vals = flat.gather(2, lin[:, None, :].expand(B, C, -1))
        return vals * (w * ok.to(w.dtype))[:, None, :]

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


def upsample_bilinear_align(plane: torch.Tensor, size: int) -> torch.Tensor:
    """F.interpolate(plane, size=size, mode='bilinear', align_corners=True).

    Stands in for pixel.py:107,110.  ``plane`` (B, C, h, w) -> (B, C, size, size).
    Source coordinate ``dst * (in - 1) / (out - 1)`` (0 when out == 1), second tap
    clamped to the last row/column (ATen UpSampleBilinear2d semantics).
    """
    B, C, h, w = plane.shape

    def axis(n_in, n_out):
        scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        dst = torch.arange(n_out, dtype=plane.dtype)
        src = dst * torch.tensor(scale, dtype=plane.dtype)
        i0 = src.floor().long().clamp(max=n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        l1 = src - i0.to(plane.dtype)
        return i0, i1, 1.0 - l1, l1

    y0, y1, hy0, hy1 = axis(h, size)
    x0, x1, wx0, wx1 = axis(w, size)
    rows0 = plane[:, :, y0, :]
    rows1 = plane[:, :, y1, :]
    top = rows0[:, :, :, x0] * wx0 + rows0[:, :, :, x1] * wx1
    bot = rows1[:, :, :, x0] * wx0 + rows1[:, :, :, x1] * wx1
    return top * hy0[:, None] + bot * hy1[:, None]
