from .resnet import resnet101  # noqa: F401
