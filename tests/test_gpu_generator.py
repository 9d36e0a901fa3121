"""On-device scene nDSM generation (generator.py) vs the CPU restatement of the reference pipeline."""
import pytest
import torch

import oracle
from oracle.generator import oracle_generate_dsm
from cases import CASES, make_cfg

pytestmark = pytest.mark.gpu


def test_scene_generation_matches_oracle():
    import tomosar2height_b200 as t2h
    from tomosar2height_b200.generator import SceneGenerator
    torch.backends.cudnn.allow_tf32 = False
    spec = CASES["berlin_small"]
    cfg = make_cfg(**spec["cfg"])  # output_size 128 -> 128 m tiles at 1 m pixels
    params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=3)
    model = t2h.TomoSAR2Height(cfg)
    model.load_state_dict(params)
    model = model.cuda()
    # a 300 m x 260 m scene in "UTM-like" coordinates, one empty corner, points exactly on tile borders
    g = torch.Generator().manual_seed(9)
    lo = torch.tensor([386000.0, 5820000.0], dtype=torch.float64)
    size = torch.tensor([300.0, 260.0], dtype=torch.float64)
    pts = torch.rand(60000, 3, generator=g, dtype=torch.float64)
    pts[:, :2] = lo + pts[:, :2] * size
    pts[:, 2] = 30.0 + pts[:, 2] * 40.0
    pts = pts[~((pts[:, 0] > lo[0] + 230) & (pts[:, 1] > lo[1] + 200))]
    pts[:50, 0] = lo[0] + 64.0   # on an anchor line: excluded by the strict crop of the tile starting there
    gen = SceneGenerator(model, lo.tolist(), (lo + size).tolist(), cfg.dataset.normalize.z_bound,
                         patch_size=128.0, stride=64.0, pixel_size=1.0, tiles_per_batch=3)
    dsm, weight = gen.generate(pts.cuda())
    ref, ref_w = oracle_generate_dsm(params, cfg, pts, lo.tolist(), (lo + size).tolist(), patch=128.0, stride=64.0, px=1.0)
    assert dsm.shape == ref.shape == (260, 300)
    assert torch.allclose(weight.cpu(), ref_w, rtol=1e-12, atol=0)      # same tiles, same windows
    covered = ref_w > 0
    scale = ref[covered].abs().max()
    assert (dsm.cpu()[covered] - ref[covered]).abs().max() <= 1e-4 * scale
    # tile-sharded generation (no collective): partial sums of two ranks add up to the same raster
    from tomosar2height_b200.parallel import shard_tiles
    parts = [gen.generate(pts.cuda(), tile_range=shard_tiles(len(gen.anchors), rk, 2)) for rk in range(2)]
    balanced = [gen.generate(pts.cuda(), rank=rk, world=3) for rk in range(3)]   # blocks balanced by candidate points
    merged3 = gen.finalize(sum(p[0] for p in balanced), sum(p[1] for p in balanced))
    assert torch.allclose(sum(p[1] for p in balanced).cpu(), ref_w, rtol=1e-12, atol=0)
    assert (merged3[covered.cuda()] - dsm[covered.cuda()]).abs().max().item() <= 1e-5 * scale.item()
    merged = gen.finalize(parts[0][0] + parts[1][0], parts[0][1] + parts[1][1])
    # not bitwise: the fp16x3 GEMMs scale their operands by the maximum of the whole batch, so a tile's result
    # depends (at fp32 rounding level) on which other tiles share its batch
    assert (merged[covered.cuda()] - dsm[covered.cuda()]).abs().max().item() <= 1e-5 * scale.item()


def test_crop_tiles_matches_oracle():
    """t2h_tile_count / t2h_tile_write against the pinned CPU restatement of crop_pc_2d + normalisation + re-crop:
    same survivors per tile (strict inequalities, points exactly on a tile edge), coordinates to float32 rounding
    (the kernel evaluates (p - min) / patch directly, the reference the algebraically equal 4x4 matrix product)."""
    import numpy as np
    import os
    from conftest import GOLDEN
    from oracle.generator import crop_normalize_tile
    from tomosar2height_b200.generator import SceneGenerator
    v = np.load(os.path.join(GOLDEN, "scene_vectors.npz"))
    pts = torch.from_numpy(v["crop_points"])
    lo, hi = [386000.0, 5820000.0], [386300.0, 5820260.0]
    gen = SceneGenerator(None, lo, hi, (-33.7, 156.5), patch_size=128.0, stride=64.0, pixel_size=1.0)
    flat, counts = gen.crop_tiles(pts.cuda(), gen.anchors)
    assert flat.shape == (sum(counts), 4) and float(flat[:, 3].abs().max()) == 0.0
    start = 0
    for (x0, y0), n in zip(gen.anchors, counts):
        idx, norm = crop_normalize_tile(pts, x0, y0, 128.0, 156.5 + 33.7)
        assert n == (0 if norm is None else norm.shape[0]), (x0, y0)
        if n:
            got = flat[start:start + n, :3].cpu().double()
            order_g = np.lexsort((got[:, 2].numpy(), got[:, 1].numpy(), got[:, 0].numpy()))
            ref = norm.double()
            order_r = np.lexsort((ref[:, 2].numpy(), ref[:, 1].numpy(), ref[:, 0].numpy()))
            assert (got[order_g] - ref[order_r]).abs().max().item() <= 2.0 ** -22   # a few float32 ulps of values < 1
        start += n
    # the two golden tiles come from the REAL reference code path
    for k in range(2):
        x0, y0 = (float(c) for c in v[f"crop_anchor_{k}"])
        t = gen.anchors.index((x0, y0)) if (x0, y0) in gen.anchors else None
        if t is not None:
            assert counts[t] == v[f"crop_norm_{k}"].shape[0]


def test_blend_accumulate_matches_reference_bitwise():
    """t2h_blend_accumulate against DSMGenerator.generate_dsm itself (golden scene from the real reference):
    flip, float64 window product and accumulation in tile order are bit-identical, also across batch boundaries."""
    import numpy as np
    import os
    from conftest import GOLDEN
    from tomosar2height_b200 import _lib
    from tomosar2height_b200.generator import SceneGenerator, blend_vectors
    v = np.load(os.path.join(GOLDEN, "scene_vectors.npz"))
    tiles = torch.from_numpy(v["scene_tiles"]).cuda()
    anchors = [tuple(float(c) for c in a) for a in v["scene_anchors"]]
    gen = SceneGenerator(None, [386000.0, 5820000.0], [386300.0, 5820260.0], (-33.7, 156.5), patch_size=128.0,
                         stride=64.0, pixel_size=1.0)
    S = 128
    wx, wy = blend_vectors(S, S, device="cuda")
    for per_batch in (1, 4, 6):
        dsm = torch.zeros(gen.n_rows, gen.n_cols, dtype=torch.float64, device="cuda")
        weight = torch.zeros_like(dsm)
        for k in range(0, len(anchors), per_batch):
            win = [gen.raster_window(*a) for a in anchors[k:k + per_batch]]
            rows, cols = [w[0] for w in win], [w[1] for w in win]
            t_rows = torch.tensor(rows, dtype=torch.int32, device="cuda")
            l_cols = torch.tensor(cols, dtype=torch.int32, device="cuda")
            _lib.call("t2h_blend_accumulate", _lib.ptr(tiles[k:k + per_batch].contiguous()), len(win), S, _lib.ptr(t_rows),
                      _lib.ptr(l_cols), _lib.ptr(wx), _lib.ptr(wy), min(rows), min(cols), max(rows) + S - min(rows),
                      max(cols) + S - min(cols), gen.n_rows, gen.n_cols, _lib.ptr(dsm), _lib.ptr(weight))
        got = gen.finalize(dsm, weight).cpu().numpy()
        ref = v["scene_dsm"]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)]), per_batch
