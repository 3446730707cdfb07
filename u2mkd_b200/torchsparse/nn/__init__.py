"""torchsparse.nn: Conv3d / BatchNorm / ReLU modules + functional + utils."""
from . import functional, utils
from .modules import BatchNorm, Conv3d, ReLU

__all__ = ["Conv3d", "BatchNorm", "ReLU", "functional", "utils"]
