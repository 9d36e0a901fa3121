"""CPU restatement of the reference's scene inference (test infrastructure; see oracle/__init__.py).

Follows dataset.py:160-181 (regular anchors), utils/crop_cloud.py:21-29 (strict crop over the whole
cloud, per tile), dataset.py:185-196,243-278 (normalisation matrix with the local z minimum, float32 cast,
re-crop), generator.py:85-113 (blend window) and generator.py:127-157 (flip, weighted accumulation, division,
clamp) with RasterData index arithmetic (utils/io_raster.py:56-62,78-95,123-131).

PINNED: every piece below is held bit-exact (raster indices, windows, crop indices, accumulated scene) or to
float32 rounding (normalised coordinates) to outputs of the REAL reference functions, produced by
tests/golden/make_golden_scene.py and committed as tests/golden/scene_vectors.npz
(tests/test_oracle_golden.py::test_scene_pieces_match_reference).
"""
import math

import numpy as np
import torch

from .model import oracle_forward


def blend_window(n_rows, n_cols, half_blend=(0.5, 0.5), min_weight=1e-3):
    """DSMGenerator._linear_blend_patch_weight (generator.py:85-113): float64 (n_rows, n_cols).
    NB the reference sizes the COLUMN ramp from n_rows and the ROW ramp from n_cols (idx_x / idx_y)."""
    wx = torch.ones(n_rows, n_cols, dtype=torch.float64)
    wy = torch.ones(n_rows, n_cols, dtype=torch.float64)
    ix, iy = math.floor(n_rows * half_blend[0]), math.floor(n_cols * half_blend[1])
    if ix > 0:
        wx[:, :ix] = torch.linspace(min_weight, 1, ix, dtype=torch.float64)[None, :]
        wx[:, -ix:] = torch.linspace(1, min_weight, ix, dtype=torch.float64)[None, :]
    if iy > 0:
        wy[:iy, :] = torch.linspace(min_weight, 1, iy, dtype=torch.float64)[:, None]
        wy[-iy:, :] = torch.linspace(1, min_weight, iy, dtype=torch.float64)[:, None]
    return wx * wy


def crop_normalize_tile(pts, x0, y0, patch, z_scale):
    """One tile of dataset.py:__getitem__ without augmentation.  ``pts`` (P, 3) float64 world coordinates.
    Returns (indices into pts, normalised float32 (n, 3)) or (empty, None) for an empty tile.

    crop_pc_2d (strict inequalities) -> z_shift = min z -> transform_mat = diag(patch, patch, z_scale, 1) with
    translation (tile centre, z_shift) -> normalize_mat = shift_norm @ inverse(transform_mat) (torch.inverse, as
    utils/coordinate.py:125-139) -> homogeneous matmul + division (apply_transform, :103-122) -> .float() ->
    re-crop strictly inside the unit square (dataset.py:278)."""
    keep = torch.where((pts[:, 0] > x0) & (pts[:, 0] < x0 + patch) & (pts[:, 1] > y0) & (pts[:, 1] < y0 + patch))[0]
    tile = pts[keep]
    if tile.shape[0] == 0:
        return keep, None
    z_shift = tile[:, 2].min()
    transform = torch.diag(torch.tensor([patch, patch, z_scale, 1.0], dtype=torch.float64))
    # translation = (tile centre, z_shift), the centre computed as (min + max) / 2 like dataset.py:266
    transform[0, 3] = (x0 + (x0 + patch)) / 2.0
    transform[1, 3] = (y0 + (y0 + patch)) / 2.0
    transform[2, 3] = z_shift
    shift_norm = torch.cat([torch.eye(4, 3, dtype=torch.float64),
                            torch.tensor([0.5, 0.5, 0.0, 1.0], dtype=torch.float64).reshape(-1, 1)], 1)
    normalize = shift_norm @ torch.eye(4, dtype=torch.float64) @ torch.eye(4, dtype=torch.float64) @ torch.inverse(transform)
    hom = torch.cat([tile, torch.ones((tile.shape[0], 1), dtype=torch.float64)], dim=1).T
    p2 = torch.matmul(normalize, hom).T
    norm = (p2[:, :3] / p2[:, 3:4]).float()
    inside = torch.where((norm[:, 0] > 0) & (norm[:, 0] < 1) & (norm[:, 1] > 0) & (norm[:, 1] < 1))[0]
    return keep[inside], norm[inside]


def raster_col_row(x, y, left, top, px):
    """RasterData.query_col_row (io_raster.py:56-62,123-131): T = Affine(px, 0, left, 0, -px, top); floor(~T * (x, y))."""
    a, c, e, f = px, left, -px, top
    ia, ic, ie, i_f = 1.0 / a, -c / a, 1.0 / e, -f / e
    return int(np.floor(x * ia + ic)), int(np.floor(y * ie + i_f))


def accumulate_scene(tiles, anchors, scene_min, scene_max, patch, px, half_blend=(0.5, 0.5)):
    """generator.py:127-157: ``tiles`` list of (S, S, 1) height grids (None = invalid tile, skipped)."""
    l, b = float(scene_min[0]), float(scene_min[1])
    r, t = float(scene_max[0]), float(scene_max[1])
    n_rows, n_cols = math.floor((t - b) / px), math.floor((r - l) / px)
    n = int(round(patch / px))
    window = blend_window(n, n, half_blend)
    dsm = torch.zeros(n_rows, n_cols, dtype=torch.float64)
    weight = torch.zeros_like(dsm)
    for h, (x0, y0) in zip(tiles, anchors):
        if h is None:
            continue
        h_grid = h[None].flip(1).squeeze()
        l_col, b_row = raster_col_row(x0 + px / 2.0, y0 + px / 2.0, l, t, px)
        r_col, t_row = raster_col_row(x0 + patch - px / 2.0, y0 + patch - px / 2.0, l, t, px)
        dsm[t_row:b_row + 1, l_col:r_col + 1] += h_grid * window
        weight[t_row:b_row + 1, l_col:r_col + 1] += window
    return dsm, weight


def oracle_generate_dsm(P, cfg, points, scene_min, scene_max, patch=512.0, stride=256.0, px=1.0, half_blend=(0.5, 0.5)):
    pts = points.double().cpu()
    l, b = float(scene_min[0]), float(scene_min[1])
    r, t = float(scene_max[0]), float(scene_max[1])
    z_bound = cfg["dataset"]["normalize"]["z_bound"]
    xs = np.concatenate([np.arange(l, r - patch, stride), [r - patch]])
    ys = np.concatenate([np.arange(b, t - patch, stride), [t - patch]])
    tiles, anchors = [], []
    for y0 in ys:
        for x0 in xs:
            _, norm = crop_normalize_tile(pts, float(x0), float(y0), patch, z_bound[1] - z_bound[0])
            anchors.append((float(x0), float(y0)))
            if norm is None or norm.shape[0] == 0:
                tiles.append(None)
                continue
            with torch.no_grad():
                tiles.append(oracle_forward(P, cfg, norm[None])[0][0])
    dsm, weight = accumulate_scene(tiles, anchors, scene_min, scene_max, patch, px, half_blend)
    dsm = torch.maximum(dsm / weight, torch.tensor(0., dtype=torch.float64))
    return dsm, weight
