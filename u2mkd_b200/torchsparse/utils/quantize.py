"""torchsparse.utils.quantize [TS v1.4.0]; host-side NumPy, runs in DataLoader workers
(core/datasets/semantic_nusc.py:326)."""
from itertools import repeat

import numpy as np

__all__ = ["sparse_quantize", "ravel_hash"]


def ravel_hash(x: np.ndarray) -> np.ndarray:
    assert x.ndim == 2, x.shape
    x = (x - np.min(x, axis=0)).astype(np.uint64, copy=False)
    extent = np.max(x, axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h = (h + x[:, k]) * extent[k + 1]
    return h + x[:, -1]


def sparse_quantize(coords, voxel_size=1, *, return_index: bool = False, return_inverse: bool = False):
    """First point of every voxel, voxels ordered by ravel hash (lexicographic x, y, z)."""
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(repeat(voxel_size, 3))
    coords = np.floor(coords / np.array(voxel_size)).astype(np.int32)
    _, indices, inverse = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
    outputs = [coords[indices]]
    if return_index:
        outputs.append(indices)
    if return_inverse:
        outputs.append(inverse)
    return outputs[0] if len(outputs) == 1 else outputs
