"""coordinate2index (reference: utils/coordinate.py:12-28) without the open3d import."""
import torch

from .. import functional as T


def coordinate2index(x, reso, coord_type='2d'):
    """Cell id ``ix + reso * iy`` of points normalised to (0, 1): (B, N, 2) -> (B, 1, N) int64.

    Bit-exact with the reference: fp32 multiply, truncation, no clamp.  CUDA tensors go through
    t2h_cell_index; anything else is rejected (no CPU fallback on the product path).
    """
    if coord_type != '2d':
        raise ValueError(f"Unsupported coord_type: {coord_type}")
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("coordinate2index: expected a CUDA tensor (the B200 path has no CPU fallback)")
    return T.cell_index(x.float() if x.dtype != torch.float32 else x, reso)
