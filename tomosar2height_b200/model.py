"""TomoSAR2Height model shell (reference: tomosar2height/model.py:8-86).

Same constructor (a cfg read by item and by attribute), same ``forward`` / ``encode_inputs``
signatures, same ``state_dict`` names and shapes, same initialisation rule.
"""
import torch
import torch.nn as nn
from typing import Dict

from .decoder import decoder_dict
from .encoder import encoder_dict


class TomoSAR2Height(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cfg_model = cfg['model']
        self.dim = cfg_model['data_dim']
        self.use_cloud = cfg.use_cloud
        self.use_image = cfg.use_image

        if self.use_cloud:
            self.point_encoder = encoder_dict[cfg_model['encoder']](dim=self.dim, **cfg_model['encoder_kwargs'])
        if self.use_image:
            self.image_encoder = encoder_dict[cfg_model.get('encoder2')](**cfg_model.get('encoder2_kwargs', {}))
        self.decoder = decoder_dict['pixel'](**cfg_model['decoder_pixel_kwargs'])

        self.threshold = cfg['test']['threshold']
        z_bound = cfg['dataset']['normalize']['z_bound']
        self.z_scale = z_bound[1] - z_bound[0]
        self._initialize_weights()
        # plane CNNs run NHWC in cuDNN; values / shapes / names of the parameters are unchanged
        self.to(memory_format=torch.channels_last)

    def _initialize_weights(self):
        """Xavier-uniform for every Conv2d / Linear weight, zero bias (model.py:46-52).  As in the
        reference, ConvTranspose2d is not a Conv2d subclass and keeps PyTorch's default init."""
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, input_cloud=None, input_image=None):
        """-> (heights (B, S, S, 1) * z_scale, footprint logits (B, S, S, 1) or None)"""
        assert self.use_image or self.use_cloud, "At least one input modality must be used."
        feature_planes = self.encode_inputs(input_cloud, input_image)
        pa, pb = self.decoder(feature_planes)
        return pa * self.z_scale, pb

    def encode_inputs(self, input_cloud=None, input_image=None):
        feature_planes = {}
        if self.use_cloud:
            cloud_features: Dict = self.point_encoder(input_cloud)
            feature_planes.update(cloud_features)
        if self.use_image:
            feature_planes['image'] = self.image_encoder(input_image)
        return feature_planes
