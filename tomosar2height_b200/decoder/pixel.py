"""PixelwiseDecoder (reference: tomosar2height/decoder/pixel.py:8-125).

The plane is bilinearly up-sampled to the output raster (t2h_upsample_bilinear_*, channels-last,
align_corners=True), the image plane is added, and a conv head (implicit GEMM, conv.py) or an FC head of
ResnetBlockFC blocks produces heights (+ optional footprint logits).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as T
from ..conv import apply_conv
from ..block import ResnetBlockFC
from ..linear import linear, relu as relu_bounded, leaky_relu as leaky_relu_bounded


class ConvDecoder(nn.Module):
    def __init__(self, in_channels=32, out_channels=1, leaky=False):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(128, 64, kernel_size=3, padding=1)
        self.conv4 = nn.Conv2d(288, out_channels, kernel_size=1)
        self.act = leaky_relu_bounded if leaky else relu_bounded  # F.leaky_relu / F.relu, handing the operand maximum on

    def forward(self, x):
        x1 = self.act(apply_conv(self.conv1, x))
        x2 = self.act(apply_conv(self.conv2, x1))
        x3 = self.act(apply_conv(self.conv3, x2))
        if not x.is_cuda:
            return apply_conv(self.conv4, torch.cat([x, x1, x2, x3], dim=1))
        # conv4 (1x1) over the dense-skip concatenation (pixel.py:31) = sum of its four column blocks applied to the
        # four tensors: two GEMM launches with two K sources each, the second taking the first as residual -- the
        # 288-channel concatenation at 512^2 (1.2 GB per 4 tiles, plus its gradient and the slicing copies in the
        # backward) is never materialised
        w = self.conv4.weight.reshape(self.conv4.weight.shape[0], -1)
        c0, c1, c2 = x.shape[1], x.shape[1] + x1.shape[1], x.shape[1] + x1.shape[1] + x2.shape[1]
        nhwc = lambda t: t.permute(0, 2, 3, 1)
        y = linear(nhwc(x), w[:, :c1], self.conv4.bias, x2=nhwc(x1))
        y = linear(nhwc(x2), w[:, c1:], None, x2=nhwc(x3), residual=y)
        return y.permute(0, 3, 1, 2)


class FCDecoder(nn.Module):
    def __init__(self, in_channels=32, out_channels=1, n_blocks=5, leaky=False):
        super().__init__()
        self.blocks = nn.ModuleList([ResnetBlockFC(in_channels) for _ in range(n_blocks)])
        self.fc_out = nn.Linear(in_channels, out_channels)
        self.act = F.leaky_relu if leaky else F.relu

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        if self.act is F.relu:
            return linear(x, self.fc_out.weight, self.fc_out.bias, relu_in=True)
        return linear(self.act(x), self.fc_out.weight, self.fc_out.bias)


class PixelwiseDecoder(nn.Module):
    def __init__(self, hidden_dim=32, out_dim=1, output_size=512, leaky=False, sample_mode='bilinear',
                 mode='conv', use_footprint=False, **kwargs):
        super().__init__()
        self.mode, self.use_footprint = mode, use_footprint
        self.sample_mode, self.output_size = sample_mode, output_size
        if mode == 'conv':
            self.conv_decoder = ConvDecoder(hidden_dim, out_dim, leaky)
            if use_footprint:
                self.conv_decoder_footprint = ConvDecoder(hidden_dim, out_dim)
        elif mode == 'fc':
            # pixel.py:88 passes `leaky` positionally into n_blocks (False -> 0 blocks, True -> 1);
            # kept verbatim so parameter names / shapes match reference checkpoints.
            self.fc_decoder = FCDecoder(hidden_dim, out_dim, leaky)
            if use_footprint:
                self.fc_decoder_footprint = FCDecoder(hidden_dim, out_dim)
        else:
            raise ValueError("Invalid mode. Use 'conv' or 'fc'.")

    def _resize(self, plane):
        """logical (B, C, h, w) -> channels-last (B, S, S, C)"""
        if self.sample_mode != 'bilinear':
            out = F.interpolate(plane, size=self.output_size, mode=self.sample_mode)
            return out.permute(0, 2, 3, 1).contiguous()
        return T.upsample_bilinear(T.nchw_to_plane(plane), self.output_size)

    def forward(self, feature_planes):
        c = None
        if 'xy' in feature_planes:
            c = self._resize(feature_planes['xy'])
        if 'image' in feature_planes:
            img = self._resize(feature_planes['image'])
            c = img if c is None else c + img
        x_footprint = None
        if self.mode == 'conv':
            c_nchw = c.permute(0, 3, 1, 2)  # channels_last strides: cuDNN runs NHWC
            x = self.conv_decoder(c_nchw).permute(0, 2, 3, 1)
            if self.use_footprint:
                x_footprint = self.conv_decoder_footprint(c_nchw).permute(0, 2, 3, 1)
        else:
            x = self.fc_decoder(c)
            if self.use_footprint:
                x_footprint = self.fc_decoder_footprint(c)
        return x, x_footprint
