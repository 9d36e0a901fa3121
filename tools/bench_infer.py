"""Config 4 (BASELINE.json): full-scene nDSM inference on a synthetic Munich-density scene.

python tools/bench_infer.py [--scale 0.45] [--gpus-emulated N]
Munich configuration (ALTO depth 6, footprint head, z range 134 m), 512 m tiles at 256 m stride, 1 m
pixels, ~1.93 points / m^2 (100 M points over 9050 m x 5730 m; --scale shrinks the scene edge lengths,
keeping the density).  Forward only, tiles batched as ragged clouds, everything on the device.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import oracle
import tomosar2height_b200 as t2h
from tomosar2height_b200.generator import SceneGenerator

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.45)
ap.add_argument("--tiles-per-batch", type=int, default=4)
args = ap.parse_args()
torch.backends.cudnn.allow_tf32 = False
cfg = t2h.munich_config()
params = oracle.synth_state_dict(oracle.reference_param_shapes(cfg), seed=0)
model = t2h.TomoSAR2Height(cfg); model.load_state_dict(params); model = model.cuda().eval()
W, H = 9050.0 * args.scale, 5730.0 * args.scale
n_pts = int(1.93 * W * H)
g = torch.Generator(device="cuda").manual_seed(0)
pts = torch.rand(n_pts, 3, generator=g, device="cuda", dtype=torch.float64)
# 70 % of the points on line-like "facades"
n_c = int(0.7 * n_pts); n_seg = max(int(W * H / 6500), 1)
a = torch.rand(n_seg, 2, generator=g, device="cuda", dtype=torch.float64)
d = (torch.rand(n_seg, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5) * (100.0 / max(W, H))
which = torch.randint(0, n_seg, (n_c,), generator=g, device="cuda")
tt = torch.rand(n_c, 1, generator=g, device="cuda", dtype=torch.float64)
pts[:n_c, :2] = (a[which] + tt * d[which] + torch.randn(n_c, 2, generator=g, device="cuda", dtype=torch.float64) * (1.0 / max(W, H))).clamp(0, 1)
lo = torch.tensor([686167.0, 5331627.0], dtype=torch.float64, device="cuda")
pts[:, 0] = lo[0] + pts[:, 0] * W; pts[:, 1] = lo[1] + pts[:, 1] * H; pts[:, 2] = 465.5 + pts[:, 2] * 60.0
gen = SceneGenerator(model, lo.tolist(), [lo[0].item() + W, lo[1].item() + H], cfg.dataset.normalize.z_bound,
                     tiles_per_batch=args.tiles_per_batch)
warm = SceneGenerator(model, lo.tolist(), [lo[0].item() + 1024, lo[1].item() + 1024], cfg.dataset.normalize.z_bound, tiles_per_batch=args.tiles_per_batch)
warm.generate(pts)  # warm-up on a corner of the scene
torch.cuda.synchronize(); t0 = time.perf_counter()
dsm, weight = gen.generate(pts)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
evals = 0.0
print(json.dumps({"config": "munich cloud-only inference, depth 6 + footprint head, 512 m tiles @ 256 m stride", "scene_m": [W, H],
                  "scene_points": n_pts, "tiles": len(gen.anchors), "seconds": round(dt, 3),
                  "scene_points_per_s": round(n_pts / dt), "ndsm_px_per_s": round(dsm.numel() / dt),
                  "tile_px_per_s": round(len(gen.anchors) * 512 * 512 / dt),
                  "covered_fraction": float((weight > 0).double().mean()), "n_gpus": 1}))
