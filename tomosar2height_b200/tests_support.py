"""Small host-side helpers shared by tests and benchmarks (no oracle dependency)."""
import torch


def _compact1by1(v):
    v = v & 0x55555555
    v = (v | (v >> 1)) & 0x33333333
    v = (v | (v >> 2)) & 0x0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF
    return v


def demorton(code: torch.Tensor):
    """Morton code (x in the even bits) -> (ix, iy), int64 tensors."""
    code = code.long()
    return _compact1by1(code), _compact1by1(code >> 1)
