"""CPU restatement of the reference's scene inference (test infrastructure; see oracle/__init__.py).

Follows dataset.py:160-181 (regular anchors), utils/crop_cloud.py:21-29 (strict crop over the whole
cloud, per tile), dataset.py:243-278 (normalisation with the local z minimum, float32 cast, re-crop),
generator.py:85-113 (blend window) and generator.py:127-157 (flip, weighted accumulation, division,
clamp) with RasterData index arithmetic (utils/io_raster.py:56-62,78-95,123-131).
"""
import math

import numpy as np
import torch

from .model import oracle_forward


def oracle_generate_dsm(P, cfg, points, scene_min, scene_max, patch=512.0, stride=256.0, px=1.0, half_blend=(0.5, 0.5)):
    pts = points.double().cpu()
    l, b = float(scene_min[0]), float(scene_min[1])
    r, t = float(scene_max[0]), float(scene_max[1])
    z_bound = cfg["dataset"]["normalize"]["z_bound"]
    n_rows, n_cols = math.floor((t - b) / px), math.floor((r - l) / px)
    xs = np.concatenate([np.arange(l, r - patch, stride), [r - patch]])
    ys = np.concatenate([np.arange(b, t - patch, stride), [t - patch]])
    n = int(round(patch / px))
    wx = torch.ones(n, n, dtype=torch.float64)
    wy = torch.ones(n, n, dtype=torch.float64)
    ix, iy = math.floor(n * half_blend[0]), math.floor(n * half_blend[1])
    if ix > 0:
        wx[:, :ix] = torch.linspace(1e-3, 1, ix, dtype=torch.float64)[None, :]
        wx[:, -ix:] = torch.linspace(1, 1e-3, ix, dtype=torch.float64)[None, :]
    if iy > 0:
        wy[:iy, :] = torch.linspace(1e-3, 1, iy, dtype=torch.float64)[:, None]
        wy[-iy:, :] = torch.linspace(1, 1e-3, iy, dtype=torch.float64)[:, None]
    window = wx * wy
    dsm = torch.zeros(n_rows, n_cols, dtype=torch.float64)
    weight = torch.zeros_like(dsm)
    for y0 in ys:
        for x0 in xs:
            keep = (pts[:, 0] > x0) & (pts[:, 0] < x0 + patch) & (pts[:, 1] > y0) & (pts[:, 1] < y0 + patch)
            tile = pts[keep]
            if tile.shape[0] == 0:
                continue
            z_shift = tile[:, 2].min()
            norm = torch.stack([(tile[:, 0] - x0) / patch, (tile[:, 1] - y0) / patch,
                                (tile[:, 2] - z_shift) / (z_bound[1] - z_bound[0])], 1).float()
            inside = (norm[:, 0] > 0) & (norm[:, 0] < 1) & (norm[:, 1] > 0) & (norm[:, 1] < 1)
            norm = norm[inside]
            if norm.shape[0] == 0:
                continue
            with torch.no_grad():
                h = oracle_forward(P, cfg, norm[None])[0]
            h_grid = h.flip(1).squeeze().double()
            l_col = math.floor((x0 + px / 2 - l) / px)
            r_col = math.floor((x0 + patch - px / 2 - l) / px)
            b_row = math.floor((t - (y0 + px / 2)) / px)
            t_row = math.floor((t - (y0 + patch - px / 2)) / px)
            dsm[t_row:b_row + 1, l_col:r_col + 1] += h_grid * window
            weight[t_row:b_row + 1, l_col:r_col + 1] += window
    dsm = torch.maximum(dsm / weight, torch.tensor(0., dtype=torch.float64))
    return dsm, weight
