"""Reduced configurations + seeded synthetic inputs shared by the golden generator and the tests.

The full Berlin / Munich configurations (conf/model/tomosar2height.yaml, conf/dataset/*.yaml)
produce 10-74 M parameters and 512x512 outputs -- too large to commit as fixtures -- so the
golden cases keep every structural feature (ALTO depth >= 3 with pooled / un-pooled levels,
the depth-2 "no-up" block, both decoders, both scatter types, the footprint head, the image
branch) at small width / resolution.
"""
import math

import numpy as np
import torch


class Cfg(dict):
    """dict with attribute access, like the OmegaConf DictConfig the reference reads
    both ways (model.py:18-21)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc


def _wrap(node):
    if isinstance(node, dict):
        return Cfg({k: _wrap(v) for k, v in node.items()})
    return node


def make_cfg(*, use_image=False, use_footprint=False, hidden=32, feat=32, reso=64, scatter="max",
             unet_type="alto", depth=4, start=32, mode="conv", leaky=False, output_size=128,
             img_depth=3, img_start=8, z_bound=(-33.7, 156.5)):
    return _wrap({
        "use_cloud": True, "use_image": use_image, "use_footprint": use_footprint, "gpu_id": 0,
        "model": {
            "name": "tomosar2height", "encoder": "pointnet_local_pool",
            "encoder_kwargs": {"hidden_dim": hidden, "feature_dim": feat, "plane_resolution": reso,
                               "scatter_type": scatter, "unet_type": unet_type,
                               "unet_kwargs": {"depth": depth, "merge_mode": "concat", "start_filts": start}},
            "encoder2": "unet",
            "encoder2_kwargs": {"num_classes": feat, "in_channels": 3, "depth": img_depth,
                                "merge_mode": "concat", "start_filts": img_start},
            "decoder_pixel_kwargs": {"mode": mode, "use_footprint": use_footprint, "hidden_dim": feat,
                                     "out_dim": 1, "sample_mode": "bilinear", "leaky": leaky,
                                     "output_size": output_size},
            "data_dim": 3,
        },
        "test": {"threshold": 0.5},
        "dataset": {"normalize": {"z_bound": list(z_bound)}},
    })


CASES = {
    # Berlin-like: cloud only, conv decoder, scatter max, ALTO depth 4 (levels 64,64,32,16)
    "berlin_small": dict(cfg=dict(), B=1, N=3000, seed=11),
    # Munich-like: cloud + image, footprint head, ALTO depth 5 (64,64,32,16,8), batch 2
    "munich_small": dict(cfg=dict(use_image=True, use_footprint=True, depth=5, start=16,
                                  z_bound=(465.5, 599.5)), B=2, N=1500, seed=12),
    # fc decoder with the pixel.py:88 quirk (leaky=True -> one block), 5-block footprint head, scatter mean
    "fc_leaky": dict(cfg=dict(use_footprint=True, reso=32, depth=3, mode="fc", leaky=True, scatter="mean",
                              output_size=64), B=1, N=1000, seed=13),
    # plane U-Net instead of ALTO (unet_type: unet)
    "plain_unet": dict(cfg=dict(unet_type="unet", reso=32, depth=3, output_size=64), B=1, N=800, seed=14),
}


def synthetic_cloud(B, N, seed, clustered=True):
    """Berlin-shaped tile: xy in the open unit square, z in [0, ~0.5].

    70 % of the points lie on ~40 random line segments ("facades", skewed cell occupancy incl.
    exact duplicates and empty cells), the rest are uniform.
    """
    g = torch.Generator().manual_seed(seed)
    lo, hi = 2.0 ** -24, 1.0 - 2.0 ** -24
    pts = torch.rand(B, N, 3, generator=g)
    if clustered:
        n_c = int(0.7 * N)
        n_seg = 40
        a = torch.rand(B, n_seg, 2, generator=g)
        d = (torch.rand(B, n_seg, 2, generator=g) - 0.5) * 0.2
        which = torch.randint(0, n_seg, (B, n_c), generator=g)
        t = torch.rand(B, n_c, 1, generator=g)
        base = torch.gather(a, 1, which[..., None].expand(-1, -1, 2))
        dirs = torch.gather(d, 1, which[..., None].expand(-1, -1, 2))
        pts[:, :n_c, :2] = base + t * dirs + 0.002 * torch.randn(B, n_c, 2, generator=g)
        # exact duplicates => argmax ties
        pts[:, 1:n_c:7] = pts[:, 0:n_c - 1:7][:, : pts[:, 1:n_c:7].shape[1]]
    pts[..., :2] = pts[..., :2].clamp(lo, hi)
    pts[..., 2] = pts[..., 2] * 0.5
    perm = torch.randperm(N, generator=g)
    return pts[:, perm].contiguous().float()


def synthetic_targets(B, size, seed, with_image=False, img_size=None):
    g = torch.Generator().manual_seed(seed + 1000)
    dsm = torch.rand(B, size, size, generator=g) * 30.0
    dsm[dsm < 6.0] = 0.0  # ground pixels => both footprint classes present
    image = torch.randn(B, 3, img_size or size, img_size or size, generator=g) if with_image else None
    return dsm, image


def grad_probe_positions(numel, k=8):
    """Fixed pseudo-random flat positions at which gradient samples are stored."""
    return [(i * 2654435761 + 12345) % numel for i in range(min(k, numel))]
