"""Drop-in for the `torchsparse` (v1.4.0) surface U2MKD uses.

`u2mkd_b200.install_as_torchsparse()` registers this package as `torchsparse` in
sys.modules, so the reference's core/models/*.py import it unchanged.  Every op runs in
libu2mkd_b200.so on CUDA tensors; there is no CPU path.
"""
from .operators import cat
from .tensor import PointTensor, SparseTensor
from . import nn, utils
from .utils import collate, quantize  # noqa: F401

__version__ = "1.4.0+u2mkd_b200"
__all__ = ["SparseTensor", "PointTensor", "cat", "nn", "utils"]
