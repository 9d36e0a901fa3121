from .resnet import ResnetBlockFC  # noqa: F401
