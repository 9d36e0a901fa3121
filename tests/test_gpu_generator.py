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
    merged = gen.finalize(parts[0][0] + parts[1][0], parts[0][1] + parts[1][1])
    # not bitwise: the fp16x3 GEMMs scale their operands by the maximum of the whole batch, so a tile's result
    # depends (at fp32 rounding level) on which other tiles share its batch
    assert (merged[covered.cuda()] - dsm[covered.cuda()]).abs().max().item() <= 1e-5 * scale.item()
