#!/usr/bin/env python
"""Generate the golden fixtures by running the REAL reference model on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference cannot be imported as-is: ``open3d``, ``rasterio``,
``laspy`` and ``torch_scatter`` are absent.  The first three are IO-only and get
empty ``sys.modules`` stubs.  ``torch_scatter`` gets the naive stand-in below: plain
python loops that follow torch_scatter 2.1.x's documented CPU rules (strict ``>``
update so ties keep the first index, empty segment -> 0 / arg = N, mean divides by
max(count, 1)).  It is deliberately written differently from oracle/ops.py
(sequential loops vs. vectorised scatter_reduce) so that the two check each other.

Nothing here is copied into the product; the outputs are small ``.npz`` / ``.json``
fixtures committed under tests/golden/.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REFERENCE = os.environ.get("T2H_REFERENCE", "/root/reference")

from cases import CASES, make_cfg, synthetic_cloud, synthetic_targets, grad_probe_positions  # noqa: E402
from oracle.model import synth_state_dict  # noqa: E402  (parameter recipe only)


# ---------------------------------------------------------------- naive torch_scatter stand-in
class _NaiveScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim_size):
        B, C, N = src.shape
        s = src.detach().numpy()
        idx = index.expand(B, 1, N).numpy()
        out = np.full((B, C, dim_size), -np.inf, dtype=s.dtype)
        arg = np.full((B, C, dim_size), N, dtype=np.int64)
        for b in range(B):
            for n in range(N):
                m = idx[b, 0, n]
                better = s[b, :, n] > out[b, :, m]
                out[b, better, m] = s[b, better, n]
                arg[b, better, m] = n
        out[arg == N] = 0
        arg_t = torch.from_numpy(arg)
        ctx.save_for_backward(arg_t)
        ctx.n = N
        ctx.mark_non_differentiable(arg_t)
        return torch.from_numpy(out), arg_t

    @staticmethod
    def backward(ctx, g_out, _g_arg):
        (arg,) = ctx.saved_tensors
        B, C, M = arg.shape
        g = np.zeros((B, C, ctx.n + 1), dtype=g_out.numpy().dtype)
        a, go = arg.numpy(), g_out.numpy()
        for b in range(B):
            for c in range(C):
                np.add.at(g[b, c], a[b, c], go[b, c])
        return torch.from_numpy(g[:, :, : ctx.n].copy()), None, None


class _NaiveScatterMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim_size):
        B, C, N = src.shape
        s = src.detach().numpy()
        idx = index.expand(B, 1, N).numpy()
        out = np.zeros((B, C, dim_size), dtype=s.dtype)
        cnt = np.zeros((B, dim_size), dtype=s.dtype)
        for b in range(B):
            for n in range(N):
                out[b, :, idx[b, 0, n]] += s[b, :, n]
                cnt[b, idx[b, 0, n]] += 1
        cnt[cnt < 1] = 1
        out /= cnt[:, None, :]
        ctx.save_for_backward(index, torch.from_numpy(cnt))
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, g_out):
        index, cnt = ctx.saved_tensors
        B, C, M = g_out.shape
        scaled = g_out / cnt[:, None, :]
        return scaled.gather(2, index.expand(B, C, -1)), None, None


def _scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert dim in (-1, 2) and out is None
    return _NaiveScatterMax.apply(src.contiguous(), index, dim_size)


def _scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    assert dim in (-1, 2)
    if out is not None:
        dim_size = out.shape[-1]
    return _NaiveScatterMean.apply(src.contiguous(), index, dim_size)


def import_reference():
    class _Anything(types.ModuleType):
        """IO-only dependency stub: any attribute (used in type hints) resolves to another stub."""

        def __getattr__(self, attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything(self.__name__ + "." + attr)

    for name in ("open3d", "rasterio", "rasterio.transform", "rasterio.crs", "rasterio.io", "laspy"):
        sys.modules.setdefault(name, _Anything(name))
    ts = types.ModuleType("torch_scatter")
    ts.scatter_max, ts.scatter_mean = _scatter_max, _scatter_mean
    sys.modules["torch_scatter"] = ts
    sys.path.insert(0, REFERENCE)
    import tomosar2height  # noqa: F401  (the reference package)
    from tomosar2height import TomoSAR2Height
    from utils.coordinate import coordinate2index
    return TomoSAR2Height, coordinate2index


def run_case(name, spec, TomoSAR2Height, coordinate2index):
    cfg = make_cfg(**spec["cfg"])
    B, N, seed = spec["B"], spec["N"], spec["seed"]
    size = cfg.model.decoder_pixel_kwargs.output_size
    cloud = synthetic_cloud(B, N, seed)
    dsm, image = synthetic_targets(B, size, seed, with_image=cfg.use_image)
    torch.manual_seed(0)
    model = TomoSAR2Height(cfg)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, f"state_dict_{name}.json"), "w") as fh:
        json.dump({k: list(v) for k, v in shapes.items()}, fh, indent=0, sort_keys=True)
    params = synth_state_dict(shapes, seed=seed)
    model.load_state_dict(params)
    out = {"cloud": cloud.numpy(), "dsm": dsm.numpy()}
    if image is not None:
        out["image"] = image.numpy()
    out["index"] = coordinate2index(cloud[:, :, :2].clone(), cfg.model.encoder_kwargs.plane_resolution).numpy()
    # fp32 only: alto.py:93 casts the sampling grid with .float(), so the reference cannot run in fp64
    for tag, dt in (("f32", torch.float32),):
        model = model.to(dt)
        model.zero_grad()
        img = image.to(dt) if image is not None else None
        pa, pb = model(input_cloud=cloud.to(dt), input_image=img)
        # trainer.py:63-69
        loss = torch.nn.functional.l1_loss(pa.squeeze(), dsm.squeeze().to(dt))
        if cfg.use_footprint:
            loss = loss + 10.0 * torch.nn.functional.binary_cross_entropy_with_logits(
                pb.squeeze(), (dsm.squeeze() > 0.0001).to(dt))
        loss.backward()
        out[f"pa_{tag}"] = pa.detach().numpy()
        if pb is not None:
            out[f"pb_{tag}"] = pb.detach().numpy()
        out[f"loss_{tag}"] = np.asarray(loss.item())
        names, norms, probes = [], [], []
        for pname, p in sorted(model.named_parameters()):
            names.append(pname)
            if p.grad is None:  # e.g. the last UpConv's unused upconv / fc_comm / fc_c
                norms.append(-1.0)
                probes.append(np.zeros(8))
                continue
            g = p.grad.detach().double().flatten()
            norms.append(g.norm().item())
            vals = [g[i].item() for i in grad_probe_positions(g.numel())]
            probes.append(np.asarray(vals + [0.0] * (8 - len(vals))))
        out[f"grad_norm_{tag}"] = np.asarray(norms)
        out[f"grad_probe_{tag}"] = np.stack(probes)
    out["param_names"] = np.asarray(names)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: loss32={out['loss_f32']:.6f} params={len(shapes)}")


def op_vectors(coordinate2index):
    """G-1 .. G-5 of SURVEY §8c: small known-answer vectors at the operator boundary."""
    import torch.nn.functional as F
    vec = {}
    # G-1: the reference's only self-check, pointnet.py:114-123
    xy = torch.tensor([[[0., 0.], [0.3, 0.9], [0.9, 0.3], [0.9, 0.9], [0.1, 0.2]]])
    idx = coordinate2index(xy, 2)
    plane = _scatter_mean(xy.permute(0, 2, 1), idx, out=xy.new_zeros(1, 2, 4)).reshape(1, 2, 2, 2)
    vec["g1_xy"], vec["g1_index"], vec["g1_plane"] = xy.numpy(), idx.numpy(), plane.numpy()
    # G-2: ties + empty segments
    src = torch.tensor([[[1., 3., 3., -2.]]])
    sidx = torch.tensor([[[0, 1, 1, 3]]])
    o, a = _scatter_max(src, sidx, dim_size=5)
    vec["g2_src"], vec["g2_index"], vec["g2_out"], vec["g2_arg"] = src.numpy(), sidx.numpy(), o.numpy(), a.numpy()
    # G-3: border points
    edge = torch.tensor([[[1 - 2.0 ** -24, 2.0 ** -24], [0.5, 0.5], [0.99999, 0.00001]]], dtype=torch.float32)
    for r in (256, 100):
        vec[f"g3_index_{r}"] = coordinate2index(edge, r).numpy()
    vec["g3_xy"] = edge.numpy()
    # G-4: grid_sample through the reference's own call (alto.py:90-95)
    plane = torch.arange(16, dtype=torch.float32).reshape(1, 1, 4, 4)
    p = torch.tensor([[[0.9, 0.1, 0.0], [0.0001, 0.9999, 0.0], [0.5, 0.5, 0.0]]])
    grid = 2.0 * p[..., [0, 1]][:, :, None].float() - 1.0
    vec["g4_plane"], vec["g4_p"] = plane.numpy(), p.numpy()
    vec["g4_out"] = F.grid_sample(plane, grid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1).numpy()
    # G-5: identity / 2x interpolate (pixel.py:107)
    g = torch.Generator().manual_seed(5)
    t = torch.rand(1, 2, 6, 6, generator=g)
    vec["g5_in"] = t.numpy()
    vec["g5_same"] = F.interpolate(t, size=6, mode="bilinear", align_corners=True).numpy()
    vec["g5_up"] = F.interpolate(t, size=12, mode="bilinear", align_corners=True).numpy()
    np.savez_compressed(os.path.join(HERE, "op_vectors.npz"), **vec)
    print("op vectors:", {k: v.shape for k, v in vec.items()})


def full_state_dicts(TomoSAR2Height):
    """Parameter names / shapes of the full Berlin and Munich configurations."""
    sys.path.insert(0, os.path.join(ROOT))
    import yaml
    from cases import _wrap
    base = yaml.safe_load(open(os.path.join(REFERENCE, "conf/model/tomosar2height.yaml")))["model"]
    for tag, depth, foot, zb in (("berlin", 5, False, [-33.7, 156.5]), ("munich", 6, True, [465.5, 599.5])):
        for use_image in (False, True):
            m = json.loads(json.dumps(base))
            m["encoder_kwargs"]["unet_kwargs"]["depth"] = depth
            m["decoder_pixel_kwargs"]["use_footprint"] = foot
            cfg = _wrap({"use_cloud": True, "use_image": use_image, "use_footprint": foot, "model": m,
                         "test": {"threshold": 0.5}, "dataset": {"normalize": {"z_bound": zb}}})
            model = TomoSAR2Height(cfg)
            shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
            fn = f"state_dict_{tag}{'_image' if use_image else ''}_full.json"
            with open(os.path.join(HERE, fn), "w") as fh:
                json.dump(shapes, fh, indent=0, sort_keys=True)
            print(fn, sum(int(np.prod(s)) for s in shapes.values()), "parameters")


if __name__ == "__main__":
    Model, c2i = import_reference()
    op_vectors(c2i)
    full_state_dicts(Model)
    for case, spec in CASES.items():
        run_case(case, spec, Model, c2i)
