from . import unet, pointnet, alto  # noqa: F401


def _out_of_scope(name):
    def build(*args, **kwargs):
        raise NotImplementedError(
            f"encoder '{name}' is outside the B200 hot path (SURVEY.md §2.1 rows 9-10); "
            "use the reference implementation for it")
    return build


# same keys as the reference registry (tomosar2height/encoder/__init__.py:3-8)
encoder_dict = {
    'pointnet_local_pool': pointnet.LocalPoolPointnet,
    'pointnet_plus_plus': _out_of_scope('pointnet_plus_plus'),
    'hourglass': _out_of_scope('hourglass'),
    'unet': unet.UNet,
}
