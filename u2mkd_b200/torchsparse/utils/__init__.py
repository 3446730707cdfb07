"""torchsparse.utils [TS v1.4.0 torchsparse/utils/utils.py]."""
from itertools import repeat
from typing import Tuple

__all__ = ["make_ntuple"]


def make_ntuple(x, ndim: int) -> Tuple[int, ...]:
    if isinstance(x, int):
        x = tuple(repeat(x, ndim))
    elif isinstance(x, list):
        x = tuple(x)
    assert isinstance(x, tuple) and len(x) == ndim, x
    return x

