"""spnn.Conv3d / BatchNorm / ReLU [TS v1.4.0 nn/modules/{conv,norm,activation}.py];
constructed at core/models/build_blocks.py:25-31,43-49,59-77 and core/models/semantickitti/spvcnn.py:31-34."""
import math

import numpy as np
import torch
from torch import nn

from ..tensor import SparseTensor
from ..utils import make_ntuple
from . import functional as F
from .utils import fapply

__all__ = ["Conv3d", "BatchNorm", "ReLU"]


class Conv3d(nn.Module):
    """Parameter `kernel` is [K, Cin, Cout] ([Cin, Cout] for a 1x1x1 kernel), K indexed in
    get_kernel_offsets order, so torchsparse-trained checkpoints load unchanged."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size=3, stride=1, dilation: int = 1,
                 bias: bool = False, transposed: bool = False) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = make_ntuple(kernel_size, ndim=3)
        self.stride = make_ntuple(stride, ndim=3)
        self.dilation = dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self) -> str:
        s = "{in_channels}, {out_channels}, kernel_size={kernel_size}"
        if self.stride != (1,) * len(self.stride):
            s += ", stride={stride}"
        if self.dilation != 1:
            s += ", dilation={dilation}"
        if self.bias is None:
            s += ", bias=False"
        if self.transposed:
            s += ", transposed=True"
        return s.format(**self.__dict__)

    def reset_parameters(self) -> None:
        fan = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        std = 1.0 / math.sqrt(fan)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, input: SparseTensor, residual=None, relu=None, want_alias=False) -> SparseTensor:
        """`residual` / `relu` / `want_alias` are used by u2mkd_b200.fusion only (ResidualBlock tail folded into this
        conv's BatchNorm epilogue, shortcut gradient folded into its dgrad); the torchsparse call signature is
        forward(input)."""
        epilogue = getattr(self, "_u2_epilogue", None)
        if epilogue is not None and relu is not None:
            epilogue = (epilogue[0], bool(relu))
        return F.conv3d(input, self.kernel, kernel_size=self.kernel_size, bias=self.bias, stride=self.stride,
                        dilation=self.dilation, transposed=self.transposed, epilogue=epilogue, residual=residual,
                        want_alias=want_alias)


class BatchNorm(nn.BatchNorm1d):
    """BatchNorm1d over SparseTensor.F; training mode runs the kernels of csrc/norm.cu."""

    def forward(self, input: SparseTensor) -> SparseTensor:
        from ... import ops
        return fapply(input, lambda f: ops.batch_norm_relu(f, self))


class ReLU(nn.ReLU):
    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
