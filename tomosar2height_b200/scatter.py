"""torch_scatter-compatible entry points backed by the sorted segmented kernels.

Signatures follow the two calls the reference makes (pointnet.py:54-56,95,109; alto.py:85,194):
``scatter_max(src, index, dim=-1, dim_size=M) -> (out, arg)`` and
``scatter_mean(src, index, dim=-1, out=None, dim_size=None) -> out`` with ``src`` (B, C, N) and
``index`` (B, 1, N) int64.  Results follow torch_scatter's CPU rules (first index wins ties,
empty segment -> 0 / arg = N) and are deterministic.
"""
import torch

from . import functional as T
from .topology import IndexLevel


def _next_supported(c):
    for s in T.SUPPORTED_C:
        if s >= c:
            return s
    raise RuntimeError(f"scatter: channel count {c} exceeds the supported maximum {T.SUPPORTED_C[-1]}")


def _to_rows(src):
    """(B, C, N) -> contiguous (B*N, Cp) rows, zero-padded to a supported channel count."""
    if not src.is_cuda or src.dtype != torch.float32 or src.dim() != 3:
        raise RuntimeError("scatter: expected a (B, C, N) float32 CUDA tensor (no CPU fallback)")
    B, C, N = src.shape
    rows = src.permute(0, 2, 1)
    Cp = _next_supported(C)
    if Cp != C:
        rows = torch.nn.functional.pad(rows, (0, Cp - C))
    return rows.reshape(B * N, Cp), C


def _check_dim(src, dim):
    if dim not in (-1, src.dim() - 1):
        raise ValueError("scatter: only reduction over the last dimension is supported")


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    _check_dim(src, dim)
    if out is not None:
        raise ValueError("scatter_max: out= is not supported")
    B, C, N = src.shape
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    level = IndexLevel(index, dim_size)
    rows, C = _to_rows(src)
    plane, arg = T.seg_max(rows, level)
    plane = plane.view(B, dim_size, -1)[:, :, :C].permute(0, 2, 1)
    base = (torch.arange(B, device=src.device, dtype=torch.int32) * N).view(B, 1, 1)
    arg = arg.view(B, dim_size, -1)[:, :, :C]
    arg64 = torch.where(arg < 0, torch.full_like(arg, N), arg - base).permute(0, 2, 1).long()
    return plane, arg64


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    _check_dim(src, dim)
    B, C, N = src.shape
    if out is not None:
        dim_size = out.shape[-1]
    elif dim_size is None:
        dim_size = int(index.max().item()) + 1
    level = IndexLevel(index, dim_size)
    rows, C = _to_rows(src)
    plane = T.seg_mean(rows, level).view(B, dim_size, -1)[:, :, :C].permute(0, 2, 1)
    if out is not None:
        # the reference only ever passes zero-initialised planes (pointnet.py:107-109)
        out.copy_(out + plane)
        return out
    return plane
