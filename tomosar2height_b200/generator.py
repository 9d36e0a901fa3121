"""Whole-scene nDSM generation on the device.

Restates the inference pipeline of the reference -- regular tile anchors (dataset.py:160-181), strict 2-D
crop (utils/crop_cloud.py:21-29), per-tile normalisation with the local z minimum (dataset.py:243-278,
conf/dataset/base.yaml:18-22), model forward, vertical flip, separable linear blend window and
weighted accumulation into float64 scene rasters (generator.py:85-113,127-157) -- with the points
resident in HBM: the scene cloud is binned once by stride-sized cells (one sort), every tile gathers its
candidates from <= 9 contiguous cell ranges instead of scanning the whole chunk on the CPU
(dataset.py:234), and tiles with different point counts are batched as a ``RaggedCloud``.
No collective is needed: with several GPUs every rank generates a block of the tile list into its own
partial rasters (``tile_range``), which are summed afterwards.
"""
import math

import torch

from .topology import RaggedCloud


def regular_anchors(lo, hi, patch, stride):
    """dataset.py:165-170: arange(lo, hi - patch, stride) plus the last anchor hi - patch."""
    out, v = [], lo
    while v < hi - patch:
        out.append(v)
        v += stride
    out.append(hi - patch)
    return out


def linear_blend_weight(n_rows, n_cols, half_blend=(0.5, 0.5), min_weight=1e-3, device="cpu"):
    """generator.py:85-113: separable ramp min_weight -> 1 -> min_weight, float64 (n_rows, n_cols)."""
    wx = torch.ones(n_rows, n_cols, dtype=torch.float64, device=device)
    wy = torch.ones(n_rows, n_cols, dtype=torch.float64, device=device)
    ix, iy = math.floor(n_rows * half_blend[0]), math.floor(n_cols * half_blend[1])
    if ix > 0:
        wx[:, :ix] = torch.linspace(min_weight, 1, ix, dtype=torch.float64, device=device)
        wx[:, -ix:] = torch.linspace(1, min_weight, ix, dtype=torch.float64, device=device)
    if iy > 0:
        wy[:iy, :] = torch.linspace(min_weight, 1, iy, dtype=torch.float64, device=device)[:, None]
        wy[-iy:, :] = torch.linspace(1, min_weight, iy, dtype=torch.float64, device=device)[:, None]
    return wx * wy


class SceneGenerator:
    def __init__(self, model, scene_min, scene_max, z_bound, patch_size=512.0, stride=256.0, pixel_size=1.0,
                 half_blend=(0.5, 0.5), tiles_per_batch=4):
        self.model = model
        self.l, self.b = float(scene_min[0]), float(scene_min[1])
        self.r, self.t = float(scene_max[0]), float(scene_max[1])
        self.patch, self.stride, self.px = float(patch_size), float(stride), float(pixel_size)
        self.z_scale = float(z_bound[1] - z_bound[0])
        self.tiles_per_batch = tiles_per_batch
        # RasterWriter.cal_dsm_shape (utils/io_raster.py:78-95)
        self.n_rows = math.floor((self.t - self.b) / self.px)
        self.n_cols = math.floor((self.r - self.l) / self.px)
        xs = regular_anchors(self.l, self.r, self.patch, self.stride)
        ys = regular_anchors(self.b, self.t, self.patch, self.stride)
        self.anchors = [(x, y) for y in ys for x in xs]  # meshgrid order of dataset.py:171-172
        self.half_blend = half_blend

    # -- binning: one sort of the scene cloud by stride-sized cell --------------------------------------
    def _bin(self, pts64):
        nx = max(int(math.ceil((self.r - self.l) / self.stride)), 1)
        ny = max(int(math.ceil((self.t - self.b) / self.stride)), 1)
        cx = ((pts64[:, 0] - self.l) / self.stride).floor().clamp(0, nx - 1).long()
        cy = ((pts64[:, 1] - self.b) / self.stride).floor().clamp(0, ny - 1).long()
        order = torch.argsort(cy * nx + cx, stable=True)
        starts = torch.searchsorted((cy * nx + cx)[order], torch.arange(nx * ny + 1, device=pts64.device))
        return pts64[order], starts.tolist(), nx, ny

    def _tile_points(self, binned, starts, nx, ny, x0, y0):
        """Strictly inside (x0, x0+patch) x (y0, y0+patch), normalised to the open unit square (fp32)."""
        x1, y1 = x0 + self.patch, y0 + self.patch
        cx0 = min(max(int(math.floor((x0 - self.l) / self.stride)), 0), nx - 1)
        cx1 = min(max(int(math.floor((x1 - self.l) / self.stride)), 0), nx - 1)
        cy0 = min(max(int(math.floor((y0 - self.b) / self.stride)), 0), ny - 1)
        cy1 = min(max(int(math.floor((y1 - self.b) / self.stride)), 0), ny - 1)
        parts = [binned[starts[cy * nx + cx0]:starts[cy * nx + cx1 + 1]] for cy in range(cy0, cy1 + 1)]
        cand = torch.cat(parts, 0) if len(parts) > 1 else parts[0]
        keep = (cand[:, 0] > x0) & (cand[:, 0] < x1) & (cand[:, 1] > y0) & (cand[:, 1] < y1)
        pts = cand[keep]
        if pts.shape[0] == 0:
            return None
        z_shift = pts[:, 2].min()  # z_shift: 'local_min'
        norm = torch.stack([(pts[:, 0] - x0) / self.patch, (pts[:, 1] - y0) / self.patch,
                            (pts[:, 2] - z_shift) / self.z_scale], 1).float()
        inside = (norm[:, 0] > 0) & (norm[:, 0] < 1) & (norm[:, 1] > 0) & (norm[:, 1] < 1)  # dataset.py:278
        norm = norm[inside]
        return norm if norm.shape[0] > 0 else None

    @torch.no_grad()
    def generate(self, points, tile_range=None):
        """points (P, 3) world coordinates on the device (float64 recommended for geo-coordinates).
        Returns (dsm (n_rows, n_cols) float64, weight float64); with ``tile_range`` the un-normalised
        partial sums of that block of tiles (sum the partials of all ranks, then ``finalize``)."""
        dev = points.device
        binned, starts, nx, ny = self._bin(points.double())
        dsm = torch.zeros(self.n_rows, self.n_cols, dtype=torch.float64, device=dev)
        weight = torch.zeros_like(dsm)
        n_px = int(round(self.patch / self.px))
        window = linear_blend_weight(n_px, n_px, self.half_blend, device=dev)
        todo = list(self.anchors if tile_range is None else [self.anchors[i] for i in tile_range])
        self.model.eval()
        for k in range(0, len(todo), self.tiles_per_batch):
            batch, clouds = [], []
            for (x0, y0) in todo[k:k + self.tiles_per_batch]:
                pts = self._tile_points(binned, starts, nx, ny, x0, y0)
                if pts is not None:  # empty tiles are skipped (generator.py:133)
                    batch.append((x0, y0))
                    clouds.append(pts)
            if not batch:
                continue
            heights = self.model(input_cloud=RaggedCloud.from_list(clouds))[0]  # (B, S, S, 1)
            for (x0, y0), h in zip(batch, heights):
                h_grid = h.flip(0).squeeze(-1).double()  # generator.py:147
                # generator.py:139-154 with RasterData.query_col_row (io_raster.py:123-131)
                l_col = math.floor((x0 + self.px / 2 - self.l) / self.px)
                r_col = math.floor((x0 + self.patch - self.px / 2 - self.l) / self.px)
                b_row = math.floor((self.t - (y0 + self.px / 2)) / self.px)
                t_row = math.floor((self.t - (y0 + self.patch - self.px / 2)) / self.px)
                dsm[t_row:b_row + 1, l_col:r_col + 1] += h_grid * window
                weight[t_row:b_row + 1, l_col:r_col + 1] += window
        if tile_range is not None:
            return dsm, weight
        return self.finalize(dsm, weight), weight

    @staticmethod
    def finalize(dsm_sum, weight_sum):
        """generator.py:156-157: divide by the accumulated weight, clamp at 0 (uncovered pixels stay NaN)."""
        return torch.maximum(dsm_sum / weight_sum, torch.zeros((), dtype=dsm_sum.dtype, device=dsm_sum.device))
