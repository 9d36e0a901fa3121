"""Per-kernel device timing and algorithmic-traffic accounting for the C-ABI calls.

``KernelTimer`` brackets every ``_lib.call`` with CUDA events on the launching stream while it is
active and accumulates, per entry point, the launch count, the device time and the ALGORITHMIC
bytes of SURVEY.md §8(d) (fp32 features, int32 indices):

  S1 fwd  seg_max_fwd      read 4NC + 4N            write 4NC (+ 4MC plane) + 4MC (arg)
  S1 bwd  seg_max_bwd      read 4NC + 4MC           write 4NC
  S2 fwd  seg_reduce_fwd   read 4NC + 4N            write 4MC
  S2 bwd  seg_broadcast    read 4MC + 4N            write 4NC
  G1 fwd  bilinear_sample_fwd  read 4MC + 8N        write 4NC
  G2 bwd  bilinear_sample_bwd  read 4NC + 8N        write 4MC
  G3      upsample fwd/bwd     read 4*B*C*h*w       write 4*B*C*H*W   (and the reverse)

N = points, M = B*r*r plane cells, C = channels.  (4N is the per-point share of the cell table /
permutation; arg is int32 here, not the reference's int64.)
"""
import torch

from . import _lib


def _bytes(name, a):
    """algorithmic bytes of one call from its positional ctypes arguments (see include/t2h.h)."""
    if name == "t2h_seg_max_fwd":   # rows, n_rows, perm, tie, keys, cell_start, n_seg, shift, C, morton, reso, ws, ws_bytes, pooled, plane, arg
        rows, n_seg, C = a[1], a[6], a[8]
        return 4 * rows * C + 4 * rows + (4 * rows * C if a[13] else 0) + (4 * n_seg * C if a[14] else 0) + 4 * n_seg * C
    if name == "t2h_seg_max_bwd":   # g_pooled, g_plane, n_rows, perm, keys, cell_start, n_seg, shift, C, ...
        rows, n_seg, C = a[2], a[6], a[8]
        return 4 * rows * C + 4 * n_seg * C + 4 * rows * C
    if name == "t2h_seg_reduce_fwd":  # rows, n_rows, perm, keys, cell_start, n_seg, shift, C, ...
        rows, n_seg, C = a[1], a[5], a[7]
        return 4 * rows * C + 4 * rows + 4 * n_seg * C
    if name == "t2h_seg_broadcast":   # plane, n_rows, perm, keys, cell_start, n_seg, shift, C, ...
        rows, n_seg, C = a[1], a[5], a[7]
        return 4 * rows * C + 4 * rows + 4 * n_seg * C
    if name == "t2h_seg_broadcast_add":  # plane, add, n_rows, perm, keys, cell_start, n_seg, shift, C, ...
        rows, n_seg, C = a[2], a[6], a[8]
        return 8 * rows * C + 4 * rows + 4 * n_seg * C
    if name == "t2h_bilinear_sample_fwd":
        reso, C, n, n_per = a[1], a[2], a[7], a[8]
        return 4 * max(n // max(n_per, 1), 1) * reso * reso * C + 8 * n + 4 * n * C
    if name == "t2h_bilinear_sample_bwd":
        rows, reso, C, n_seg = a[1], a[2], a[3], a[9]  # g, n, reso, C, xyz, stride, perm, keys, cell_start, n_seg, ...
        return 4 * rows * C + 8 * rows + 4 * n_seg * C
    if name in ("t2h_upsample_bilinear_fwd", "t2h_upsample_bilinear_bwd"):
        B, h, w, C, oh, ow = a[1:7]
        return 4 * B * C * (h * w + oh * ow)
    if name in ("t2h_gather_rows", "t2h_scatter_rows"):
        n, width = a[2], a[3]
        return 8 * n * width + 4 * n
    if name == "t2h_xy_keys":
        return 12 * a[1] + 4 * a[1]
    if name == "t2h_sort_by_cell":
        return 4 * 16 * a[1] + 4 * a[2]  # 4 radix passes x (key+value in, key+value out) + cell table
    if name == "t2h_cell_index":
        return 16 * a[1]
    if name == "t2h_linear_fwd":  # x1, ld, k1, x2, ld, k2, rows, w_hi, w_lo, n_out, bias, relu, mask, ld, res, ld, out, ld
        k, rows, n = a[2] + a[5], a[6], a[9]
        return 4 * rows * (k + n + (n if a[12] else 0) + (n if a[14] else 0)) + 8 * k * n
    if name == "t2h_linear_wgrad":  # g, ld, x, ld, rows, n_out, k_in, relu, ws, ws_bytes, gw, ld, gb
        rows, n, k = a[4], a[5], a[6]
        return 4 * rows * (n + k) + 4 * n * k
    if name == "t2h_linear_fwd_f16":  # x1, ld, k1, x2, ld, k2, rows, x_max, w_hi, w_lo, w_max, n_out, bias, relu, mask, ld, res, ld, out, ld, out_max
        k, rows, n = a[2] + a[5], a[6], a[11]
        return 4 * rows * (k + n + (n if a[14] else 0) + (n if a[16] else 0)) + 4 * k * n
    if name == "t2h_linear_wgrad_f16":  # g, ld, g_max, x, ld, x_max, rows, n_out, k_in, ...
        rows, n, k = a[6], a[7], a[8]
        return 4 * rows * (n + k) + 4 * n * k
    if name == "t2h_add_absmax":  # a, b, n, out, slot
        return 12 * a[2]
    if name == "t2h_absmax":  # x1, ld, k1, x2, ld, k2, rows, slot
        return 4 * a[6] * (a[2] + a[5])
    if name == "t2h_colsum":
        return 4 * a[2] * a[3]
    if name == "t2h_conv3x3_fwd":
        px = a[1] * a[2] * a[3]
        return 4 * px * (a[4] + a[7] + (a[7] if a[10] else 0) + (a[7] if a[11] else 0)) + 8 * 9 * a[4] * a[7]
    if name == "t2h_conv3x3_wgrad":
        px = a[2] * a[3] * a[4]
        return 4 * px * (a[5] + a[6]) + 4 * 9 * a[5] * a[6]
    if name == "t2h_conv3x3_fwd_f16":  # x, B, H, W, cin, x_max, w_hi, w_lo, w_max, cout, bias, relu, mask, res, out, out_max
        px = a[1] * a[2] * a[3]
        return 4 * px * (a[4] + a[9] + (a[9] if a[12] else 0) + (a[9] if a[13] else 0)) + 4 * 9 * a[4] * a[9]
    if name == "t2h_conv3x3_wgrad_f16":  # g, g_max, x, x_max, B, H, W, cin, cout
        px = a[4] * a[5] * a[6]
        return 4 * px * (a[7] + a[8]) + 4 * 9 * a[7] * a[8]
    return 0


def _flops(name, a):
    """algorithmic FLOPs (2*M*K*N, one pass; the three TF32 passes of the split are not counted)."""
    if name == "t2h_linear_fwd":
        return 2 * a[6] * (a[2] + a[5]) * a[9]
    if name == "t2h_linear_wgrad":
        return 2 * a[4] * a[5] * a[6]
    if name == "t2h_conv3x3_fwd_f16":
        return 2 * a[1] * a[2] * a[3] * 9 * a[4] * a[9]
    if name == "t2h_conv3x3_wgrad_f16":
        return 2 * a[4] * a[5] * a[6] * 9 * a[7] * a[8]
    if name == "t2h_linear_fwd_f16":
        return 2 * a[6] * (a[2] + a[5]) * a[11]
    if name == "t2h_linear_wgrad_f16":
        return 2 * a[6] * a[7] * a[8]
    if name == "t2h_conv3x3_fwd":   # x, B, H, W, cin, w_hi, w_lo, cout, ...
        return 2 * a[1] * a[2] * a[3] * 9 * a[4] * a[7]
    if name == "t2h_conv3x3_wgrad":  # g, x, B, H, W, cin, cout, ...
        return 2 * a[2] * a[3] * a[4] * 9 * a[5] * a[6]
    return 0


_bytes.n_rows = 0


class KernelTimer:
    """with KernelTimer(n_rows=B*N) as kt: ...   then kt.summary() -> {name: {...}}"""

    def __init__(self, n_rows):
        self.n_rows = n_rows
        self.records = {}
        self.shapes = {}
        self._orig = None

    def __enter__(self):
        self._orig = _lib.call
        _bytes.n_rows = self.n_rows
        timer = self

        def timed_call(name, *args):
            start = torch.cuda.Event(enable_timing=True)
            stop = torch.cuda.Event(enable_timing=True)
            start.record()
            timer._orig(name, *args)
            stop.record()
            timer.records.setdefault(name, []).append((start, stop, _bytes(name, args), _flops(name, args)))
            if name in ("t2h_linear_fwd", "t2h_linear_wgrad", "t2h_conv3x3_fwd", "t2h_conv3x3_wgrad",
                        "t2h_linear_fwd_f16", "t2h_linear_wgrad_f16", "t2h_conv3x3_fwd_f16", "t2h_conv3x3_wgrad_f16"):
                timer.shapes.setdefault((name, _shape(name, args)), []).append((start, stop, _flops(name, args)))

        _lib.call = timed_call
        for mod in _patch_targets():
            mod.call = timed_call
        return self

    def __exit__(self, *exc):
        _lib.call = self._orig
        for mod in _patch_targets():
            mod.call = self._orig
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = sum(r[0].elapsed_time(r[1]) for r in recs)
            nbytes = sum(r[2] for r in recs)
            flops = sum(r[3] for r in recs)
            out[name] = {
                "launches": len(recs),
                "ms_total": ms,
                "ms_avg": ms / len(recs),
                "bytes_per_launch": nbytes / len(recs),
                "gbs": (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0,
                "tflops": (flops / 1e12) / (ms / 1e3) if ms > 0 else 0.0,
            }
        return out


def _shape(name, a):
    if name == "t2h_linear_fwd":
        return (a[6], a[2] + a[5], a[9])            # rows, K, N
    if name == "t2h_linear_wgrad":
        return (a[4], a[6], a[5])                   # rows, K, N
    if name == "t2h_conv3x3_fwd_f16":
        return (a[1] * a[2] * a[3], 9 * a[4], a[9])
    if name == "t2h_conv3x3_wgrad_f16":
        return (a[4] * a[5] * a[6], 9 * a[7], a[8])
    if name == "t2h_linear_fwd_f16":  # + epilogue operand: mask (input gradient), residual, none
        return (a[6], a[2] + a[5], a[11], "mask" if a[14] else ("res" if a[16] else ""))
    if name == "t2h_linear_wgrad_f16":
        return (a[6], a[8], a[7])
    if name == "t2h_conv3x3_fwd":
        return (a[1] * a[2] * a[3], 9 * a[4], a[7])  # pixels, 9*cin, cout
    if name == "t2h_conv3x3_wgrad":
        return (a[2] * a[3] * a[4], 9 * a[5], a[6])
    return ()


def shape_table(timer):
    """[(name, (rows, K, N), launches, ms_total, TFLOP/s)] sorted by time."""
    torch.cuda.synchronize()
    out = []
    for (name, shp), recs in timer.shapes.items():
        ms = sum(r[0].elapsed_time(r[1]) for r in recs)
        fl = sum(r[2] for r in recs)
        out.append((name, shp, len(recs), ms, fl / 1e9 / ms if ms > 0 else 0.0))
    return sorted(out, key=lambda r: -r[3])


def _patch_targets():
    from . import functional
    return [functional]  # modules that bound `call` by name; the others go through _lib.call
