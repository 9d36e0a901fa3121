"""ALTO U-Net: plane CNN alternating with point <-> plane exchanges
(reference: tomosar2height/encoder/alto.py:48-382).

Per level: 2x conv3x3+ReLU on the plane (implicit GEMM on the tcgen05 pipeline, conv.py) -> residual 1x1 from the previous
level -> bilinear SAMPLE of the plane at the points (t2h_bilinear_sample_*) -> per-point
``fc_comm`` (C -> 2C -> C) + ``fc_c(c_last)`` -> cell-wise MEAN back onto the plane
(t2h_seg_reduce_*), which REPLACES the conv output.

B200 design: the points are sorted once by Morton cell code (``Topology``); every level's
segments are key ranges of that one sort, per-point features stay in sorted order for the whole
network, planes stay channels-last so point kernels write coalesced rows and cuDNN sees NHWC.
``p`` may be the reference's (B, N, 3) tensor or an already-built ``Topology``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as T
from ..conv import apply_conv
from ..linear import use_f16, publish_absmax, operand_absmax, fork2
from ..linear import linear
from ..linear import relu as relu_bounded
from ..topology import Topology
from .unet import conv3x3, conv1x1, upconv2x2, check_modes, xavier_normal_convs


def _comm_mlp(channels):
    return nn.Sequential(nn.Linear(channels, 2 * channels), nn.ReLU(), nn.Linear(2 * channels, channels))


class _Exchange:
    """sample -> fc_comm (+ fc_c) -> mean-scatter, shared by DownConv and UpConv."""

    def sample_plane_feature(self, p, c):
        """plane ``c`` (B, C, r, r) sampled at the points of topology ``p`` -> (B*N, C) sorted rows
        (alto.py:90-95 / 199-205)."""
        level = p.level(c.shape[2])
        plane = T.nchw_to_plane(c)
        sampled = T.bilinear_sample(plane, level)
        if use_f16(2 * c.shape[1], c.shape[1]):
            # operand scale of the fp16 GEMM that consumes the samples: bilinear interpolation is a convex
            # combination, so max |plane| (a small tensor, often already known) bounds max |sampled|
            publish_absmax(sampled, operand_absmax(plane.view(-1, plane.shape[-1]), owner=plane))
        return sampled

    @staticmethod
    def generate_plane_features(p, c, channel, reso_plane):
        """rows ``c`` (B*N, channel) -> mean plane (B, channel, r, r) (alto.py:76-88 / 187-197)."""
        level = p.level(reso_plane)
        return T.plane_to_nchw(T.seg_mean(c, level), p.B, reso_plane)

    def exchange(self, p, plane, c_last):
        sampled = self.sample_plane_feature(p, plane)
        hidden = linear(sampled, self.fc_comm[0].weight, self.fc_comm[0].bias)
        carry = None if c_last is None else linear(c_last, self.fc_c.weight, self.fc_c.bias)
        c = linear(hidden, self.fc_comm[2].weight, self.fc_comm[2].bias, relu_in=True, residual=carry)
        # c feeds the mean-scatter AND the next level's fc_c: the gather of the plane gradient, the sum of the two
        # gradient branches and the operand maximum of the GEMMs that consume it are one kernel in the backward
        level = p.level(plane.shape[2])
        scattered, c_next = T.seg_mean_carry(c, level)
        return T.plane_to_nchw(scattered, p.B, plane.shape[2]), c_next


class DownConv(nn.Module, _Exchange):
    def __init__(self, in_channels, out_channels, i, pooling, depth, sample_mode='bilinear'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.pooling, self.downsample, self.depth = pooling, i, depth
        self.sample_mode = sample_mode
        self.conv1 = conv3x3(in_channels, out_channels)
        self.conv2 = conv3x3(out_channels, out_channels)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.fc_comm = _comm_mlp(out_channels)
        self.fc_c = nn.Linear(in_channels, out_channels)
        if i > 0:
            self.conv1x1 = conv1x1(in_channels, out_channels)

    def forward(self, p, x, x_after_conv=None, c_last=None):
        # conv3x3 -> ReLU -> conv3x3 -> ReLU; the first ReLU is applied on load by the second conv
        plane = relu_bounded(apply_conv(self.conv2, apply_conv(self.conv1, x['xy']), relu_in=True))
        if x_after_conv is not None:
            side = x_after_conv['xy']
            if 2 <= self.downsample < self.depth:  # alto.py:108-110: levels >= 2 see a pooled residual
                side = self.pool(side)
            plane = apply_conv(self.conv1x1, side, residual=plane)  # plane + conv1x1(side), added in the epilogue
        x_after_conv = {'xy': plane}
        scattered, c = self.exchange(p, plane, c_last)
        before_pool = {'xy': scattered}
        x = {'xy': self.pool(scattered) if self.pooling else scattered}
        return x, before_pool, x_after_conv, c


class UpConv(nn.Module, _Exchange):
    def __init__(self, in_channels, out_channels, i, depth, merge_mode='concat', up_mode='transpose',
                 sample_mode='bilinear'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.merge_mode, self.up_mode, self.depth = merge_mode, up_mode, depth
        self.sample_mode = sample_mode
        last = i == depth - 2
        self.upconv = upconv2x2(in_channels, out_channels, mode=up_mode)
        if last:
            self.upconv_noup = conv1x1(in_channels, out_channels)
        self.fc_comm = _comm_mlp(out_channels)
        self.fc_c = nn.Linear(in_channels, out_channels)
        self.conv1x1 = conv1x1(in_channels, out_channels) if last else upconv2x2(in_channels, out_channels, mode=up_mode)
        self.conv1 = conv3x3(2 * out_channels if merge_mode == 'concat' else out_channels, out_channels)
        self.conv2 = conv3x3(out_channels, out_channels)

    def forward(self, p, from_down, from_up, x_after_conv, c_last, i):
        last = i == self.depth - 2
        up = apply_conv(self.upconv_noup, from_up['xy']) if last else apply_conv(self.upconv, from_up['xy'])
        merged = torch.cat((up, from_down['xy']), 1) if self.merge_mode == 'concat' else up + from_down['xy']
        plane = relu_bounded(apply_conv(self.conv2, apply_conv(self.conv1, merged), relu_in=True))
        if x_after_conv is not None:
            plane = apply_conv(self.conv1x1, x_after_conv['xy'], residual=plane)
        x_after_conv = {'xy': plane}
        if last:  # alto.py:241-242: the last block has no point exchange
            return {'xy': plane}, x_after_conv, c_last
        scattered, c = self.exchange(p, plane, c_last)
        return {'xy': scattered}, x_after_conv, c


class UNet(nn.Module):
    def __init__(self, num_classes, in_channels=3, depth=0, start_filts=64, up_mode='transpose',
                 merge_mode='concat', **kwargs):
        super().__init__()
        check_modes(up_mode, merge_mode)
        self.up_mode, self.merge_mode = up_mode, merge_mode
        self.num_classes, self.in_channels = num_classes, in_channels
        self.start_filts, self.depth = start_filts, depth
        self.down_convs = nn.ModuleList()
        self.up_convs = nn.ModuleList()
        outs = in_channels
        for i in range(depth):
            ins, outs = outs, start_filts * (2 ** i)
            self.down_convs.append(DownConv(ins, outs, i, pooling=0 < i < depth - 1, depth=depth))
        for i in range(depth - 1):
            ins, outs = outs, outs // 2
            self.up_convs.append(UpConv(ins, outs, i, up_mode=up_mode, merge_mode=merge_mode, depth=depth))
        self.conv_final = conv1x1(outs, num_classes)
        xavier_normal_convs(self)

    def forward(self, p, x, c):
        """p: (B, N, 3) points or a Topology; x: {'xy': (B, C, r, r)}; c: per-point features,
        (B, N, C) in point order when p is a tensor, (B*N, C) sorted rows when p is a Topology."""
        if not isinstance(p, Topology):
            topo = Topology(p, x['xy'].shape[2])
            if c is not None:
                c = topo.sort_rows(c.reshape(-1, c.shape[-1]))
            p = topo
        skips, x_after_conv = [], None
        for down in self.down_convs:
            x, before_pool, x_after_conv, c = down(p, x, x_after_conv, c)
            skips.append(before_pool)
        for i, up in enumerate(self.up_convs):
            x, x_after_conv, c = up(p, skips[-(i + 2)], x, x_after_conv, c, i)
        return apply_conv(self.conv_final, x['xy'])
