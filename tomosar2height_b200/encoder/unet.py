"""Plain convolutional U-Net (reference: tomosar2height/encoder/unet.py:48-187).

Used as the image encoder (``encoder2: unet``) and as the non-ALTO plane network
(``unet_type: unet``).  It stays on stock PyTorch / cuDNN convolutions (north star: the image CNN
is not re-written); only the parameter names / shapes are contractual.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def conv3x3(in_channels, out_channels, stride=1, padding=1, bias=True, groups=1):
    return nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=padding, bias=bias, groups=groups)


def conv1x1(in_channels, out_channels, groups=1):
    return nn.Conv2d(in_channels, out_channels, kernel_size=1, groups=groups, stride=1)


def upconv2x2(in_channels, out_channels, mode='transpose'):
    if mode == 'transpose':
        return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2)
    return nn.Sequential(nn.Upsample(mode='bilinear', scale_factor=2), conv1x1(in_channels, out_channels))


def check_modes(up_mode, merge_mode):
    """Argument validation shared by both U-Nets (unet.py:127-134, alto.py:298-319)."""
    if up_mode not in ('transpose', 'upsample'):
        raise ValueError(f"\"{up_mode}\" is not a valid mode for upsampling. Only \"transpose\" and \"upsample\" are allowed.")
    if merge_mode not in ('concat', 'add'):
        raise ValueError(f"\"{merge_mode}\" is not a valid mode for merging up and down paths. Only \"concat\" and \"add\" are allowed.")
    if up_mode == 'upsample' and merge_mode == 'add':
        raise ValueError("up_mode \"upsample\" is incompatible with merge_mode \"add\".")


def xavier_normal_convs(module):
    """reset_params of the reference (unet.py:163-171): Xavier-normal Conv2d weights, zero bias."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_normal_(m.weight)
            nn.init.constant_(m.bias, 0)


class DownConv(nn.Module):
    def __init__(self, in_channels, out_channels, pooling=True):
        super().__init__()
        self.in_channels, self.out_channels, self.pooling = in_channels, out_channels, pooling
        self.conv1 = conv3x3(in_channels, out_channels)
        self.conv2 = conv3x3(out_channels, out_channels)
        if pooling:
            self.pool = nn.MaxPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        before_pool = F.relu(self.conv2(F.relu(self.conv1(x))))
        return (self.pool(before_pool) if self.pooling else before_pool), before_pool


class UpConv(nn.Module):
    def __init__(self, in_channels, out_channels, merge_mode='concat', up_mode='transpose'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.merge_mode, self.up_mode = merge_mode, up_mode
        self.upconv = upconv2x2(in_channels, out_channels, mode=up_mode)
        self.conv1 = conv3x3(2 * out_channels if merge_mode == 'concat' else out_channels, out_channels)
        self.conv2 = conv3x3(out_channels, out_channels)

    def forward(self, from_down, from_up):
        up = self.upconv(from_up)
        x = torch.cat((up, from_down), 1) if self.merge_mode == 'concat' else up + from_down
        return F.relu(self.conv2(F.relu(self.conv1(x))))


class UNet(nn.Module):
    def __init__(self, num_classes, in_channels=3, depth=5, start_filts=64, up_mode='transpose',
                 merge_mode='concat', **kwargs):
        super().__init__()
        check_modes(up_mode, merge_mode)
        self.num_classes, self.in_channels = num_classes, in_channels
        self.start_filts, self.depth = start_filts, depth
        self.down_convs = nn.ModuleList()
        self.up_convs = nn.ModuleList()
        outs = in_channels
        for i in range(depth):
            ins, outs = outs, start_filts * (2 ** i)
            self.down_convs.append(DownConv(ins, outs, pooling=i < depth - 1))
        for _ in range(depth - 1):
            ins, outs = outs, outs // 2
            self.up_convs.append(UpConv(ins, outs, up_mode=up_mode, merge_mode=merge_mode))
        self.conv_final = conv1x1(outs, num_classes)
        xavier_normal_convs(self)

    def forward(self, x):
        skips = []
        for down in self.down_convs:
            x, before_pool = down(x)
            skips.append(before_pool)
        for i, up in enumerate(self.up_convs):
            x = up(skips[-(i + 2)], x)
        return self.conv_final(x)
