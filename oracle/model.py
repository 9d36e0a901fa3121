"""Whole-path CPU restatement of TomoSAR2Height.forward (test infrastructure).

Functional form: the network is a pure function of a ``{name: tensor}`` parameter
dictionary that uses the reference's ``state_dict`` names, so the same
dictionary can be loaded into the reference, the oracle and the B200 package.

Follows (all paths relative to /root/reference):
  model.py:54-86            forward / encode_inputs, height scaled by z_scale
  encoder/pointnet.py:60-111 LocalPoolPointnet
  block/resnet.py:46-54      ResnetBlockFC
  encoder/alto.py:97-138, 207-257, 368-382  ALTO DownConv / UpConv / UNet
  encoder/unet.py:67-73, 99-109, 173-187    plain U-Net (image branch)
  decoder/pixel.py:27-32, 54-58, 94-125     PixelwiseDecoder (+ the :88 quirk)
  trainer.py:63-69           loss
"""
import zlib

import torch
import torch.nn.functional as F

from . import ops


# --------------------------------------------------------------------------
# parameter bookkeeping
# --------------------------------------------------------------------------
def _get(cfg, *path, default=None):
    node = cfg
    for key in path:
        if node is None:
            return default
        try:
            node = node[key]
        except (KeyError, TypeError):
            return default
    return node


def reference_param_shapes(cfg) -> dict:
    """Names and shapes of the reference ``state_dict`` for ``cfg``.

    Independent restatement of the constructor wiring (model.py:18-37,
    pointnet.py:36-45, alto.py:51-74,147-182,330-354, unet.py:54-65,82-97,
    145-160, pixel.py:17-23,45-51,83-90); checked against the real reference's
    ``state_dict`` by tests/golden (state_dict_*.json).
    """
    shapes = {}

    def lin(name, n_in, n_out, bias=True):
        shapes[name + ".weight"] = (n_out, n_in)
        if bias:
            shapes[name + ".bias"] = (n_out,)

    def conv(name, n_in, n_out, k):
        shapes[name + ".weight"] = (n_out, n_in, k, k)
        shapes[name + ".bias"] = (n_out,)

    def convT(name, n_in, n_out, k=2):
        shapes[name + ".weight"] = (n_in, n_out, k, k)
        shapes[name + ".bias"] = (n_out,)

    def resblock(name, n_in, n_out=None, n_h=None):
        n_out = n_in if n_out is None else n_out
        n_h = min(n_in, n_out) if n_h is None else n_h
        lin(name + ".fc_0", n_in, n_h)
        lin(name + ".fc_1", n_h, n_out)
        if n_in != n_out:
            lin(name + ".shortcut", n_in, n_out, bias=False)

    m = cfg["model"]
    if cfg["use_cloud"]:
        ek = m["encoder_kwargs"]
        h, c_dim = ek["hidden_dim"], ek["feature_dim"]
        n_blocks = _get(ek, "n_blocks", default=5)
        pe = "point_encoder"
        lin(pe + ".fc_pos", m["data_dim"], 2 * h)
        for i in range(n_blocks):
            resblock(f"{pe}.blocks.{i}", 2 * h, h)
        lin(pe + ".fc_c", h, c_dim)
        uk = ek["unet_kwargs"]
        depth, start = uk["depth"], uk["start_filts"]
        concat = _get(uk, "merge_mode", default="concat") == "concat"
        un = pe + ".unet"
        if _get(ek, "unet_type", default="alto") == "alto":
            outs = None
            for i in range(depth):
                ins = c_dim if i == 0 else outs
                outs = start * 2 ** i
                q = f"{un}.down_convs.{i}"
                conv(q + ".conv1", ins, outs, 3)
                conv(q + ".conv2", outs, outs, 3)
                lin(q + ".fc_comm.0", outs, 2 * outs)
                lin(q + ".fc_comm.2", 2 * outs, outs)
                lin(q + ".fc_c", ins, outs)
                if i > 0:
                    conv(q + ".conv1x1", ins, outs, 1)
            for j in range(depth - 1):
                ins = outs
                outs = ins // 2
                q = f"{un}.up_convs.{j}"
                last = j == depth - 2
                convT(q + ".upconv", ins, outs)
                if last:
                    conv(q + ".upconv_noup", ins, outs, 1)
                lin(q + ".fc_comm.0", outs, 2 * outs)
                lin(q + ".fc_comm.2", 2 * outs, outs)
                lin(q + ".fc_c", ins, outs)
                if last:
                    conv(q + ".conv1x1", ins, outs, 1)
                else:
                    convT(q + ".conv1x1", ins, outs)
                conv(q + ".conv1", 2 * outs if concat else outs, outs, 3)
                conv(q + ".conv2", outs, outs, 3)
            conv(un + ".conv_final", outs, c_dim, 1)
        else:
            _plain_unet_shapes(un, c_dim, c_dim, depth, start, concat, conv, convT)
    if cfg["use_image"]:
        ik = m["encoder2_kwargs"]
        _plain_unet_shapes("image_encoder", ik["in_channels"], ik["num_classes"], ik["depth"],
                           ik["start_filts"], _get(ik, "merge_mode", default="concat") == "concat",
                           conv, convT)
    dk = m["decoder_pixel_kwargs"]
    hd, od = dk["hidden_dim"], dk["out_dim"]
    heads = ["decoder.{}_decoder"] + (["decoder.{}_decoder_footprint"] if dk["use_footprint"] else [])
    for k, head in enumerate(heads):
        if dk["mode"] == "conv":
            q = head.format("conv")
            conv(q + ".conv1", hd, 64, 3)
            conv(q + ".conv2", 64, 128, 3)
            conv(q + ".conv3", 128, 64, 3)
            conv(q + ".conv4", 288, od, 1)
        else:
            q = head.format("fc")
            # pixel.py:88 passes `leaky` positionally into n_blocks; :90 keeps the default 5
            n_fc = int(bool(dk["leaky"])) if k == 0 else 5
            for i in range(n_fc):
                resblock(f"{q}.blocks.{i}", hd)
            lin(q + ".fc_out", hd, od)
    return shapes


def _plain_unet_shapes(pre, c_in, n_cls, depth, start, concat, conv, convT):
    outs = None
    for i in range(depth):
        ins = c_in if i == 0 else outs
        outs = start * 2 ** i
        conv(f"{pre}.down_convs.{i}.conv1", ins, outs, 3)
        conv(f"{pre}.down_convs.{i}.conv2", outs, outs, 3)
    for j in range(depth - 1):
        ins = outs
        outs = ins // 2
        convT(f"{pre}.up_convs.{j}.upconv", ins, outs)
        conv(f"{pre}.up_convs.{j}.conv1", 2 * outs if concat else outs, outs, 3)
        conv(f"{pre}.up_convs.{j}.conv2", outs, outs, 3)
    conv(pre + ".conv_final", outs, n_cls, 1)


def synth_state_dict(shapes: dict, seed: int = 0, dtype=torch.float32) -> dict:
    """Deterministic, construction-order-independent parameters.

    Weights ~ U(-a, a) with the Xavier bound a = sqrt(6 / (fan_in + fan_out)),
    biases ~ U(-0.1, 0.1) (non-zero so bias paths are exercised).  Each tensor
    has its own generator seeded from (seed, crc32(name)), so any implementation
    can be given identical values through ``load_state_dict``.
    """
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        gen = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63))
        if len(shape) == 1:
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * 0.1
        else:
            rf = 1
            for s in shape[2:]:
                rf *= s
            bound = (6.0 / ((shape[0] + shape[1]) * rf)) ** 0.5
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------
# discrete selections (ReLU masks, max-pool winners, scatter-max argmax)
# --------------------------------------------------------------------------
class Selections:
    """Record the discrete selections of a forward pass, or replay recorded ones.

    The network is piecewise linear: its parameter gradients are exact functions of the inputs GIVEN the pattern of
    ReLU masks, max-pool winners and scatter-max argmax.  Two fp32 implementations that differ by rounding can pick
    different winners for near-ties, which moves single gradients by O(1e-3) although every operator is accurate to
    1e-6.  ``mode='record'`` stores the pattern of this evaluation on ``tape``; ``mode='replay'`` evaluates the
    network with the pattern on ``tape`` (e.g. the one the CUDA path took), so that gradients can be compared at
    the north-star tolerance without that ambiguity (tests/test_gpu_selection_flips.py)."""

    def __init__(self, mode="record", tape=None):
        assert mode in ("record", "replay")
        self.mode, self.tape, self.pos = mode, ([] if tape is None else tape), 0

    def _next(self, kind):
        k, v = self.tape[self.pos]
        assert k == kind, f"selection tape out of step: expected {kind}, found {k} at {self.pos}"
        self.pos += 1
        return v

    def relu(self, x, slope=0.0):
        if self.mode == "record":
            mask = x > 0
            self.tape.append(("relu", mask))
        else:
            mask = self._next("relu")
            assert mask.shape == x.shape, (mask.shape, x.shape)
        return torch.where(mask, x, x * slope)

    def max_pool(self, x):
        if self.mode == "record":
            out, idx = F.max_pool2d(x, 2, 2, return_indices=True)
            self.tape.append(("pool", idx))
            return out
        idx = self._next("pool")
        return x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)

    def seg_max(self, src, index, dim_size):
        if self.mode == "record":
            out, arg = ops.segment_max(src, index, dim_size)
            self.tape.append(("argmax", arg))
            return out, arg
        arg = self._next("argmax")
        n = src.shape[2]
        picked = src.gather(2, arg.clamp(max=n - 1))
        return torch.where(arg == n, torch.zeros((), dtype=src.dtype), picked), arg


class _Plain:
    """default policy: the operators themselves"""

    @staticmethod
    def relu(x, slope=0.0):
        return F.leaky_relu(x, slope) if slope else F.relu(x)

    @staticmethod
    def max_pool(x):
        return F.max_pool2d(x, 2, 2)

    @staticmethod
    def seg_max(src, index, dim_size):
        return ops.segment_max(src, index, dim_size)


# --------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------
def _lin(P, name, x):
    return F.linear(x, P[name + ".weight"], P.get(name + ".bias"))


def _conv(P, name, x, pad=0):
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], padding=pad)


def _convT(P, name, x):
    return F.conv_transpose2d(x, P[name + ".weight"], P[name + ".bias"], stride=2)


def _resblock(P, name, x, sel=_Plain):
    """block/resnet.py:46-54"""
    net = _lin(P, name + ".fc_0", sel.relu(x))
    dx = _lin(P, name + ".fc_1", sel.relu(net))
    if name + ".shortcut.weight" in P:
        return _lin(P, name + ".shortcut", x) + dx
    return x + dx


def _sample(plane, xy, aten):
    """plane (B,C,r,r), xy (B,N,2) -> (B,N,C); alto.py:90-95,122"""
    if aten:
        grid = (2.0 * xy - 1.0)[:, :, None]
        s = F.grid_sample(plane, grid, padding_mode="border", align_corners=True, mode="bilinear").squeeze(-1)
    else:
        s = ops.bilinear_sample_points(plane, xy)
    return s.transpose(1, 2)


def _mean_plane(c, xy, reso):
    """c (B,N,C) -> (B,C,reso,reso); alto.py:76-88 / pointnet.py:101-111"""
    idx = ops.cell_index(xy, reso)
    plane = ops.segment_mean(c.permute(0, 2, 1), idx, reso * reso)
    return plane.reshape(c.shape[0], c.shape[2], reso, reso)


def _upsample(plane, size, aten):
    if aten:
        return F.interpolate(plane, size=size, mode="bilinear", align_corners=True)
    return ops.upsample_bilinear_align(plane, size)


def _alto(P, pre, cloud, plane, c, depth, concat, aten, trace, sel=_Plain):
    """alto.py:368-382 with DownConv :97-138 and UpConv :207-257 unrolled."""
    xy = cloud[..., :2]
    x, after, skips = plane, None, []
    for i in range(depth):
        q = f"{pre}.down_convs.{i}"
        x = sel.relu(_conv(P, q + ".conv1", x, 1))
        x = sel.relu(_conv(P, q + ".conv2", x, 1))
        if after is not None:
            side = sel.max_pool(after) if 2 <= i < depth else after
            x = x + _conv(P, q + ".conv1x1", side)
        after = x
        s = _sample(x, xy, aten)
        s = _lin(P, q + ".fc_comm.2", sel.relu(_lin(P, q + ".fc_comm.0", s)))
        c = s if c is None else s + _lin(P, q + ".fc_c", c)
        x = _mean_plane(c, xy, x.shape[2])
        if trace is not None:
            trace[f"down{i}.c"] = c
            trace[f"down{i}.plane"] = x
        skips.append(x)
        if 0 < i < depth - 1:
            x = sel.max_pool(x)
    for j in range(depth - 1):
        q = f"{pre}.up_convs.{j}"
        last = j == depth - 2
        up = _conv(P, q + ".upconv_noup", x) if last else _convT(P, q + ".upconv", x)
        skip = skips[-(j + 2)]
        x = torch.cat((up, skip), 1) if concat else up + skip
        x = sel.relu(_conv(P, q + ".conv1", x, 1))
        x = sel.relu(_conv(P, q + ".conv2", x, 1))
        if after is not None:
            x = x + (_conv(P, q + ".conv1x1", after) if last else _convT(P, q + ".conv1x1", after))
        after = x
        if last:
            break
        s = _sample(x, xy, aten)
        s = _lin(P, q + ".fc_comm.2", sel.relu(_lin(P, q + ".fc_comm.0", s)))
        c = s if c is None else s + _lin(P, q + ".fc_c", c)
        x = _mean_plane(c, xy, x.shape[2])
        if trace is not None:
            trace[f"up{j}.c"] = c
            trace[f"up{j}.plane"] = x
    return _conv(P, pre + ".conv_final", x)


def _plain_unet(P, pre, x, depth, concat):
    """unet.py:173-187"""
    skips = []
    for i in range(depth):
        x = F.relu(_conv(P, f"{pre}.down_convs.{i}.conv1", x, 1))
        x = F.relu(_conv(P, f"{pre}.down_convs.{i}.conv2", x, 1))
        skips.append(x)
        if i < depth - 1:
            x = F.max_pool2d(x, 2, 2)
    for j in range(depth - 1):
        up = _convT(P, f"{pre}.up_convs.{j}.upconv", x)
        skip = skips[-(j + 2)]
        x = torch.cat((up, skip), 1) if concat else up + skip
        x = F.relu(_conv(P, f"{pre}.up_convs.{j}.conv1", x, 1))
        x = F.relu(_conv(P, f"{pre}.up_convs.{j}.conv2", x, 1))
    return _conv(P, pre + ".conv_final", x)


def _point_encoder(P, cfg, cloud, aten, trace, sel=_Plain):
    """pointnet.py:60-90"""
    ek = cfg["model"]["encoder_kwargs"]
    reso = ek["plane_resolution"]
    xy = cloud[:, :, :2]
    idx = ops.cell_index(xy, reso)
    pe = "point_encoder"
    net = _lin(P, pe + ".fc_pos", cloud)
    net = _resblock(P, pe + ".blocks.0", net, sel)
    n_blocks = sum(1 for k in P if k.startswith(pe + ".blocks.") and k.endswith(".fc_0.weight"))
    use_max = _get(ek, "scatter_type", default="max") == "max"
    for i in range(1, n_blocks):
        src = net.permute(0, 2, 1)
        if use_max:
            cells, arg = sel.seg_max(src, idx, reso * reso)
            if trace is not None:
                trace[f"pool{i}.arg"] = arg
        else:
            cells = ops.segment_mean(src, idx, reso * reso)
        pooled = cells.gather(2, idx.expand(-1, src.shape[1], -1)).permute(0, 2, 1)
        net = _resblock(P, f"{pe}.blocks.{i}", torch.cat([net, pooled], dim=2), sel)
    c = _lin(P, pe + ".fc_c", sel.relu(net))
    plane = _mean_plane(c, xy, reso)
    if trace is not None:
        trace["index"] = idx
        trace["enc.c"] = c
        trace["enc.plane"] = plane
    uk = ek["unet_kwargs"]
    concat = _get(uk, "merge_mode", default="concat") == "concat"
    if _get(ek, "unet_type", default="alto") == "alto":
        return _alto(P, pe + ".unet", cloud, plane, c, uk["depth"], concat, aten, trace, sel)
    return _plain_unet(P, pe + ".unet", plane, uk["depth"], concat)


def _decoder(P, cfg, planes, aten, output_size, sel=_Plain):
    """pixel.py:94-125"""
    dk = cfg["model"]["decoder_pixel_kwargs"]
    c = None
    if "xy" in planes:
        c = _upsample(planes["xy"], output_size, aten)
    if "image" in planes:
        img = _upsample(planes["image"], output_size, aten)
        c = img if c is None else c + img
    leaky = bool(dk["leaky"])

    def conv_head(q, slope):
        act = lambda t: sel.relu(t, slope)
        x1 = act(_conv(P, q + ".conv1", c, 1))
        x2 = act(_conv(P, q + ".conv2", x1, 1))
        x3 = act(_conv(P, q + ".conv3", x2, 1))
        return _conv(P, q + ".conv4", torch.cat([c, x1, x2, x3], 1)).permute(0, 2, 3, 1)

    def fc_head(q):
        x = c.permute(0, 2, 3, 1)
        n_fc = sum(1 for k in P if k.startswith(q + ".blocks.") and k.endswith(".fc_0.weight"))
        for i in range(n_fc):
            x = _resblock(P, f"{q}.blocks.{i}", x, sel)
        return _lin(P, q + ".fc_out", sel.relu(x))  # :88 quirk => act is always relu

    foot = None
    if dk["mode"] == "conv":
        pa = conv_head("decoder.conv_decoder", 0.01 if leaky else 0.0)  # F.leaky_relu's default slope
        if dk["use_footprint"]:
            foot = conv_head("decoder.conv_decoder_footprint", 0.0)
    elif dk["mode"] == "fc":
        pa = fc_head("decoder.fc_decoder")
        if dk["use_footprint"]:
            foot = fc_head("decoder.fc_decoder_footprint")
    else:
        raise ValueError("Invalid mode. Use 'conv' or 'fc'.")
    return pa, foot


def oracle_forward(P, cfg, input_cloud=None, input_image=None, aten=True, trace=None, selections=None):
    """model.py:54-67.  Returns (heights (B,S,S,1) * z_scale, footprint logits or None).

    ``aten=True`` uses ATen's own grid_sample / interpolate (the reference's
    dependency); ``aten=False`` uses the explicit restatements in oracle/ops.py.
    ``trace`` (dict) collects intermediate tensors for op-level parity tests.
    ``selections`` (a ``Selections``) records or replays the ReLU / max-pool / argmax pattern of the point branch
    and the decoder (the image branch keeps the plain operators).
    """
    sel = _Plain if selections is None else selections
    assert cfg["use_cloud"] or cfg["use_image"], "At least one input modality must be used."
    planes = {}
    if cfg["use_cloud"]:
        planes["xy"] = _point_encoder(P, cfg, input_cloud, aten, trace, sel)
    if cfg["use_image"]:
        ik = cfg["model"]["encoder2_kwargs"]
        planes["image"] = _plain_unet(P, "image_encoder", input_image, ik["depth"],
                                      _get(ik, "merge_mode", default="concat") == "concat")
    if trace is not None:
        trace["planes"] = dict(planes)
    output_size = _get(cfg, "model", "decoder_pixel_kwargs", "output_size", default=512)
    pa, pb = _decoder(P, cfg, planes, aten, output_size, sel)
    z_bound = cfg["dataset"]["normalize"]["z_bound"]
    return pa * (z_bound[1] - z_bound[0]), pb


def oracle_loss(pa, pb, dsm, use_footprint, weight_ce=10.0):
    """trainer.py:63-69: L1 on heights (+ weight_ce * BCE-with-logits on footprint)."""
    loss = F.l1_loss(pa.squeeze(), dsm.squeeze().to(pa.dtype))
    if use_footprint:
        target = (dsm.squeeze() > 0.0001).to(pa.dtype)
        loss = loss + weight_ce * F.binary_cross_entropy_with_logits(pb.squeeze(), target)
    return loss
