"""LocalPoolPointnet (reference: tomosar2height/encoder/pointnet.py:13-111).

fc_pos 3->2h, n_blocks ResnetBlockFC(2h, h) with cell-wise pooling + gather-back + concat between
blocks, fc_c h->C, cell-wise mean onto the (B, C, R, R) plane, then the (ALTO) U-Net.

B200 design: ``Topology`` sorts the tile batch once; the per-point pipeline runs in sorted order
(per-point MLPs are order-agnostic and only planes leave this module), pooling / mean are
deterministic segmented reductions (t2h_seg_max_*, t2h_seg_reduce_*), no index tensors are
materialised per call and no (B, C, N) transposes are made.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as T
from ..linear import linear
from .. import scatter as S
from ..block import ResnetBlockFC
from ..topology import Topology, RaggedCloud
from .unet import UNet
from .alto import UNet as Alto


class LocalPoolPointnet(nn.Module):
    def __init__(self, feature_dim=128, dim=3, hidden_dim=128, scatter_type='max', unet_type='alto',
                 unet_kwargs=None, plane_resolution=None, n_blocks=5):
        super().__init__()
        self.c_dim = feature_dim
        self.fc_pos = nn.Linear(dim, 2 * hidden_dim)
        self.blocks = nn.ModuleList([ResnetBlockFC(2 * hidden_dim, hidden_dim) for _ in range(n_blocks)])
        self.fc_c = nn.Linear(hidden_dim, feature_dim)
        self.actvn = nn.ReLU()
        self.unet_type = unet_type
        unet_kwargs = unet_kwargs or {}
        if unet_type == 'unet':
            self.unet = UNet(feature_dim, in_channels=feature_dim, **unet_kwargs)
        elif unet_type == 'alto':
            self.unet = Alto(feature_dim, in_channels=feature_dim, **unet_kwargs)
        else:
            raise ValueError(f"Unknown unet_type: {unet_type}")
        self.reso_plane = plane_resolution
        if scatter_type not in ('max', 'mean'):
            raise ValueError("Invalid scatter type")
        self.scatter_type = scatter_type
        self.scatter = S.scatter_max if scatter_type == 'max' else S.scatter_mean

    def forward(self, inputs: torch.Tensor):
        """inputs (B, N, 3) fp32 CUDA (or a RaggedCloud of tiles with different point counts), xy in the
        open unit square -> {'xy': (B, C, R, R)}."""
        if isinstance(inputs, RaggedCloud):
            topo = Topology(inputs.points, self.reso_plane, offsets=inputs.offsets)
        else:
            topo = Topology(inputs, self.reso_plane)
        level = topo.level(self.reso_plane)
        # xyz rows are stored zero-padded to 4 floats (16-byte rows for TMA); pad fc_pos.weight to match
        pad = topo.xyz_sorted.shape[1] - self.fc_pos.weight.shape[1]
        net = linear(topo.xyz_sorted, F.pad(self.fc_pos.weight, (0, pad)), self.fc_pos.bias)
        net = self.blocks[0](net)
        for block in self.blocks[1:]:
            if self.scatter_type == 'max':
                pooled = T.seg_max_pool(net, level)
            else:
                pooled = T.seg_broadcast(T.seg_mean(net, level), level)
            net = block(net, pooled)  # [net | pooled] without the concat copy
        c = linear(net, self.fc_c.weight, self.fc_c.bias, relu_in=True)
        plane = T.plane_to_nchw(T.seg_mean(c, level), topo.B, self.reso_plane)
        if self.unet_type == 'unet':
            return {'xy': self.unet(plane)}
        return {'xy': self.unet(topo, {'xy': plane}, c)}

    # -- reference-signature helpers (arbitrary point order; each call sorts by the given index) ----
    def pool_local(self, i, net):
        """i (B, 1, N) int64 cell ids, net (B, N, C) -> pooled features gathered back (B, N, C)."""
        fea = self.scatter(net.permute(0, 2, 1), i, dim_size=self.reso_plane ** 2)
        if self.scatter_type == 'max':
            fea = fea[0]
        fea = fea.gather(dim=2, index=i.expand(-1, net.size(2), -1))
        return fea.permute(0, 2, 1)

    def generate_plane_features(self, index_dict, c, plane):
        index = index_dict.get(plane)
        if index is None:
            raise NotImplementedError(f"Plane type {plane} not implemented.")
        fea_plane = S.scatter_mean(c.permute(0, 2, 1), index, dim_size=self.reso_plane ** 2)
        return fea_plane.reshape(c.size(0), self.c_dim, self.reso_plane, self.reso_plane)
