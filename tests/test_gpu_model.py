"""Whole-path GPU parity: TomoSAR2Height forward + backward on the B200 path vs
(a) fixtures produced by the real reference (tests/golden/*.npz) and (b) the CPU oracle in fp64."""
import os

import numpy as np
import pytest
import torch

import oracle
from cases import CASES, make_cfg, grad_probe_positions, synthetic_cloud, synthetic_targets

pytestmark = pytest.mark.gpu

REL = 1e-4       # north-star tolerance: heights, loss, per-operator values and gradients
GRAD_REL = 2e-3  # whole-model parameter gradients: the network is full of discrete selections (scatter-max
                 # argmax, max-pool, ReLU, sign() of the L1 loss) and ONE flipped selection among the ~1e3
                 # points of a fixture moves a gradient by ~1e-3; the CPU fp32 reference itself sits ~1e-3
                 # from an fp64 evaluation (see test_model_matches_fp64_oracle).  The claim is PROVEN in
                 # tests/test_gpu_selection_flips.py: with the selections pinned (recorded on the GPU, replayed in
                 # the oracle) every parameter gradient agrees to 1e-4, and the flips are counted


@pytest.fixture(autouse=True)
def strict_fp32():
    """Parity is defined against the CPU fp32 reference: keep the retained cuDNN convs out of TF32."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _build(name):
    import tomosar2height_b200 as t2h
    spec = CASES[name]
    cfg = make_cfg(**spec["cfg"])
    shapes = oracle.reference_param_shapes(cfg)
    params = oracle.synth_state_dict(shapes, seed=spec["seed"])
    model = t2h.TomoSAR2Height(cfg)
    model.load_state_dict(params)
    return cfg, params, model.cuda()


@pytest.mark.parametrize("name", list(CASES))
def test_model_matches_reference_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg, params, model = _build(name)
    cloud = torch.from_numpy(g["cloud"]).cuda()
    image = torch.from_numpy(g["image"]).cuda() if "image" in g.files else None
    dsm = torch.from_numpy(g["dsm"]).cuda()
    pa, pb = model(input_cloud=cloud, input_image=image)
    assert pa.shape == g["pa_f32"].shape
    scale = np.abs(g["pa_f32"]).max()
    err = np.abs(pa.detach().cpu().numpy() - g["pa_f32"]).max()
    assert err <= REL * scale, f"heights: {err:.3e} vs {scale:.3e}"
    if pb is not None:
        errb = np.abs(pb.detach().cpu().numpy() - g["pb_f32"]).max()
        assert errb <= REL * np.abs(g["pb_f32"]).max()
    # trainer.py:63-69
    loss = torch.nn.functional.l1_loss(pa.squeeze(), dsm.squeeze())
    if cfg.use_footprint:
        loss = loss + 10.0 * torch.nn.functional.binary_cross_entropy_with_logits(
            pb.squeeze(), (dsm.squeeze() > 0.0001).float())
    assert abs(loss.item() - float(g["loss_f32"])) <= REL * abs(float(g["loss_f32"]))
    loss.backward()
    named = dict(model.named_parameters())
    norm_err, probe_err = [], []
    for k, pname in enumerate(str(n) for n in g["param_names"]):
        ref_norm = float(g["grad_norm_f32"][k])
        grad = named[pname].grad
        if ref_norm < 0:
            assert grad is None or float(grad.abs().max()) == 0.0, pname
            continue
        flat = grad.double().flatten().cpu()
        norm_err.append(abs(flat.norm().item() - ref_norm) / max(ref_norm, 1e-6))
        pos = grad_probe_positions(flat.numel())
        got = np.asarray([flat[i].item() for i in pos])
        ref = g["grad_probe_f32"][k][: len(pos)]
        probe_err.append(np.abs(got - ref).max() / max(float(flat.abs().max()), 1e-12))
    # single selection flips move individual gradients by ~1e-3 (see GRAD_REL); judge the population
    assert np.median(norm_err) <= GRAD_REL and np.median(probe_err) <= GRAD_REL, (np.median(norm_err), np.median(probe_err))
    assert max(norm_err) <= 10 * GRAD_REL and max(probe_err) <= 10 * GRAD_REL, (max(norm_err), max(probe_err))


@pytest.mark.parametrize("name", ["berlin_small", "munich_small"])
def test_model_matches_fp64_oracle(name):
    """Error of the fp32 B200 path measured against an fp64 evaluation of the same network."""
    cfg, params, model = _build(name)
    spec = CASES[name]
    B, N = spec["B"], 2 * spec["N"]
    size = cfg.model.decoder_pixel_kwargs.output_size
    cloud = synthetic_cloud(B, N, seed=spec["seed"] + 50)
    dsm, image = synthetic_targets(B, size, spec["seed"] + 50, with_image=cfg.use_image)
    # A smooth functional of the outputs (fixed random weights) instead of the trainer's L1, whose
    # sign() gradient turns 1e-6 height differences into O(1) gradient flips.  Even so the network is
    # full of discrete selections (scatter-max argmax, max-pool, ReLU) that flip under fp32 rounding:
    # the CPU fp32 reference path itself is ~1e-3 away from fp64 on these gradients.  The bar is
    # therefore: heights within REL of fp64, gradients no further from fp64 than a small multiple of
    # the reference's own fp32 evaluation (op-level gradient parity at REL is in test_gpu_ops /
    # test_gpu_linear, where no selection can flip).
    g = torch.Generator().manual_seed(7)
    wa = torch.randn(B, size, size, 1, generator=g, dtype=torch.float64)
    wb = torch.randn(B, size, size, 1, generator=g, dtype=torch.float64)

    def functional(pa_, pb_, dt, dev):
        f = (pa_ * wa.to(dt).to(dev)).mean()
        return f if pb_ is None else f + (pb_ * wb.to(dt).to(dev)).mean()

    P64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    pa64, pb64 = oracle.oracle_forward(P64, cfg, cloud.double(), None if image is None else image.double(), aten=False)
    functional(pa64, pb64, torch.float64, "cpu").backward()
    P32 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pa32, pb32 = oracle.oracle_forward(P32, cfg, cloud, image, aten=False)
    functional(pa32, pb32, torch.float32, "cpu").backward()
    pa, pb = model(input_cloud=cloud.cuda(), input_image=None if image is None else image.cuda())
    functional(pa, pb, torch.float32, "cuda").backward()

    scale = pa64.abs().max().item()
    assert (pa.detach().cpu().double() - pa64.detach()).abs().max().item() <= REL * scale
    err_gpu, err_cpu = [], []
    for pname, p in model.named_parameters():
        g64 = P64[pname].grad
        if g64 is None or p.grad is None:
            continue
        denom = max(g64.abs().max().item(), 1e-12)
        err_gpu.append((p.grad.cpu().double() - g64).abs().max().item() / denom)
        err_cpu.append((P32[pname].grad.double() - g64).abs().max().item() / denom)
    med = lambda v: sorted(v)[len(v) // 2]
    assert med(err_gpu) <= 4 * med(err_cpu) + REL, (med(err_gpu), med(err_cpu))
    assert max(err_gpu) <= 4 * max(err_cpu) + 10 * REL, (max(err_gpu), max(err_cpu))


def test_alto_unet_accepts_reference_arguments():
    """UNet.forward(p, x, c) with the reference's tensor arguments (alto.py:368) == topology path."""
    cfg, params, model = _build("berlin_small")
    enc = model.point_encoder
    cloud = synthetic_cloud(1, 2000, seed=4).cuda()
    from tomosar2height_b200.topology import Topology
    import tomosar2height_b200.functional as T
    R, C = enc.reso_plane, enc.c_dim
    g = torch.Generator().manual_seed(0)
    c = torch.randn(1, 2000, C, generator=g).cuda()
    topo = Topology(cloud, R)
    plane = T.plane_to_nchw(T.seg_mean(topo.sort_rows(c.view(-1, C)), topo.level(R)), 1, R)
    with torch.no_grad():
        a = enc.unet(cloud, {'xy': plane}, c)
        b = enc.unet(topo, {'xy': plane}, topo.sort_rows(c.view(-1, C)))
    assert torch.equal(a, b)
    # pool_local / generate_plane_features with reference-style arguments
    from tomosar2height_b200.utils import coordinate2index
    idx = coordinate2index(cloud[..., :2].contiguous(), R)
    net = torch.randn(1, 2000, 32, generator=g).cuda()
    pooled = enc.pool_local(idx, net)
    cells, _ = oracle.segment_max(net.cpu().permute(0, 2, 1), idx.cpu(), R * R)
    want = cells.gather(2, idx.cpu().expand(-1, 32, -1)).permute(0, 2, 1)
    assert torch.equal(pooled.cpu(), want)
    fea = enc.generate_plane_features({'xy': idx}, c, 'xy')
    want = oracle.segment_mean(c.cpu().permute(0, 2, 1), idx.cpu(), R * R).reshape(1, C, R, R)
    assert (fea.cpu() - want).abs().max() <= 1e-5 * want.abs().max()
    with pytest.raises(NotImplementedError):
        enc.generate_plane_features({}, c, 'xz')


def test_ragged_like_batches_are_independent():
    """Tiles of one batch do not influence each other (tile-sharded data parallelism relies on it)."""
    cfg, params, model = _build("berlin_small")
    a = synthetic_cloud(1, 1500, seed=1).cuda()
    b = synthetic_cloud(1, 1500, seed=2).cuda()
    with torch.no_grad():
        both, _ = model(input_cloud=torch.cat([a, b], 0))
        ya, _ = model(input_cloud=a)
        yb, _ = model(input_cloud=b)
    assert (both[0] - ya[0]).abs().max() <= 1e-5 * ya.abs().max()
    assert (both[1] - yb[0]).abs().max() <= 1e-5 * yb.abs().max()


def test_cuda_graph_replay_matches_eager():
    """GraphedTrainStep (forward + loss + backward captured once) == the eager step, also after the
    weights have been updated in place (the TF32 weight splits must be recomputed inside the graph)."""
    from tomosar2height_b200.graph import GraphedTrainStep
    from tomosar2height_b200.parallel import FlatGradients
    cfg, params, model = _build("berlin_small")
    flat = FlatGradients(model)
    size = cfg.model.decoder_pixel_kwargs.output_size
    clouds = [synthetic_cloud(2, 3000, seed=s).cuda() for s in (21, 22)]
    dsms = [synthetic_targets(2, size, s)[0].cuda() for s in (21, 22)]

    def loss_fn(m, cloud, dsm):
        pa, _ = m(input_cloud=cloud)
        return (pa.squeeze(-1) - dsm).abs().mean(dim=(1, 2)).sum()

    graphed = GraphedTrainStep(model, loss_fn, clouds[0], dsms[0])
    for round_ in range(2):
        for cloud, dsm in zip(clouds, dsms):
            flat.zero_()
            loss_e = loss_fn(model, cloud, dsm)
            loss_e.backward()
            grads_e = flat.flat.clone()
            flat.zero_()
            loss_g = graphed(cloud, dsm).clone()
            torch.cuda.synchronize()
            assert torch.equal(loss_g, loss_e.detach()), (round_, loss_g.item(), loss_e.item())
            assert torch.equal(flat.flat, grads_e), "graph replay must be bitwise identical to the eager step"
        with torch.no_grad():  # an optimizer-style in-place update
            for p in model.parameters():
                p.mul_(1.01)


def test_ragged_batch_equals_per_tile_forward_backward():
    """Tiles with different point counts in ONE batch (flat points + offsets), incl. an empty tile,
    give the same heights and gradients as running the tiles one by one."""
    import tomosar2height_b200 as t2h
    cfg, params, model = _build("berlin_small")
    sizes = [1500, 0, 3777, 64]
    tiles = [synthetic_cloud(1, max(n, 1), seed=30 + k)[0, :n].cuda() for k, n in enumerate(sizes)]
    size = cfg.model.decoder_pixel_kwargs.output_size
    g = torch.Generator().manual_seed(5)
    w = torch.randn(len(sizes), size, size, 1, generator=g).cuda()
    w[1] = 0  # the empty tile (the reference skips such tiles: dataset.py:235-241) takes no part in the loss
    model.zero_grad()
    pa, _ = model(input_cloud=t2h.RaggedCloud.from_list(tiles))
    assert pa.shape == (len(sizes), size, size, 1) and bool(torch.isfinite(pa).all())
    (pa * w).sum().backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad()
    for k, t in enumerate(tiles):
        if t.shape[0] == 0:
            continue
        y, _ = model(input_cloud=t[None])
        (y * w[k:k + 1]).sum().backward()
        assert (pa[k].detach() - y[0].detach()).abs().max() <= 1e-5 * y.abs().max(), k
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        denom = max(float(p.grad.abs().max()), 1e-12)
        # batched and per-tile runs slice segments / split reductions differently (fp32 summation order), which
        # flips a few ReLU / max-pool selections; the mean deviation stays at rounding level
        diff = (grads[n] - p.grad).abs()
        assert float(diff.max()) <= 5e-3 * denom and float(diff.mean()) <= 1e-4 * denom, n


def test_alto_unet_tensor_arguments_carry_gradients():
    """UNet.forward(p: Tensor, x, c) -- the reference signature (alto.py:368): the gradient reaches ``c`` through the
    differentiable row permutation (t2h::gather_rows / scatter_rows) exactly as on the topology path."""
    cfg, params, model = _build("berlin_small")
    enc = model.point_encoder
    from tomosar2height_b200.topology import Topology
    import tomosar2height_b200.functional as T
    cloud = synthetic_cloud(1, 2000, seed=4).cuda()
    R, C = enc.reso_plane, enc.c_dim
    g = torch.Generator().manual_seed(0)
    c0 = torch.randn(1, 2000, C, generator=g).cuda()
    topo = Topology(cloud, R)
    plane = T.plane_to_nchw(T.seg_mean(topo.sort_rows(c0.view(-1, C)), topo.level(R)), 1, R).detach()
    w = torch.randn(1, C, R, R, generator=g).cuda()
    c_a = c0.clone().requires_grad_(True)
    (enc.unet(cloud, {'xy': plane}, c_a) * w).sum().backward()
    c_b = c0.clone().requires_grad_(True)
    (enc.unet(topo, {'xy': plane}, topo.sort_rows(c_b.view(-1, C))) * w).sum().backward()
    assert c_a.grad is not None and float(c_a.grad.abs().max()) > 0
    assert torch.equal(c_a.grad, c_b.grad)
