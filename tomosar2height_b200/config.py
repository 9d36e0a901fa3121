"""Hydra-free mirrors of the reference configuration tree (conf/config.yaml,
conf/model/tomosar2height.yaml, conf/dataset/{berlin,munich}.yaml).

Only the keys the model reads are reproduced; an OmegaConf DictConfig built by the
reference's own Hydra entry points works just as well, the model accepts either.
"""
import copy


class Config(dict):
    """Nested dict with attribute access, as model.py:18-21 reads cfg both ways."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


def to_config(node):
    if isinstance(node, dict):
        return Config({k: to_config(v) for k, v in node.items()})
    if isinstance(node, (list, tuple)):
        return [to_config(v) for v in node]
    return node


_MODEL = {  # conf/model/tomosar2height.yaml:3-30
    "name": "tomosar2height",
    "encoder": "pointnet_local_pool",
    "encoder_kwargs": {
        "hidden_dim": 32, "feature_dim": 32, "plane_resolution": 256, "scatter_type": "max",
        "unet_type": "alto",
        "unet_kwargs": {"depth": 5, "merge_mode": "concat", "start_filts": 32},
    },
    "encoder2": "unet",
    "encoder2_kwargs": {"num_classes": 32, "in_channels": 3, "depth": 6, "merge_mode": "concat", "start_filts": 32},
    "decoder_pixel_kwargs": {"mode": "conv", "use_footprint": False, "hidden_dim": 32, "out_dim": 1,
                             "sample_mode": "bilinear", "leaky": False},
    "data_dim": 3,
}


def _config(depth, use_footprint, z_bound, use_cloud, use_image):
    model = copy.deepcopy(_MODEL)
    model["encoder_kwargs"]["unet_kwargs"]["depth"] = depth
    model["decoder_pixel_kwargs"]["use_footprint"] = use_footprint
    return to_config({
        "use_cloud": use_cloud, "use_image": use_image, "use_footprint": use_footprint, "gpu_id": 0,
        "model": model,
        "training": {"batch_size": 1, "optimize_every": 64, "learning_rate": 1e-4, "weight_ce": 10.0},
        "test": {"threshold": 0.5},
        "dataset": {"normalize": {"z_bound": list(z_bound)}, "patch_size": [512, 512]},
    })


def berlin_config(use_image=False, use_cloud=True):
    """conf/dataset/berlin.yaml: ALTO depth 5, no footprint head, z_bound [-33.7, 156.5]."""
    return _config(5, False, (-33.7, 156.5), use_cloud, use_image)


def munich_config(use_image=False, use_cloud=True):
    """conf/dataset/munich.yaml:6-11,50-51: ALTO depth 6, footprint head, z_bound [465.5, 599.5]."""
    return _config(6, True, (465.5, 599.5), use_cloud, use_image)
