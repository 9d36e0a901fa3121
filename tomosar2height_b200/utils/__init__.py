from .coordinate import coordinate2index  # noqa: F401
