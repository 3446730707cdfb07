"""sptr.utils (third_party/SparseTransformer/sptr/utils.py) without torch_scatter / torch_geometric."""
import numbers

import numpy as np
import torch

from .functional import precompute_all


def to_3d_numpy(size):
    """sptr/utils.py:9-19."""
    if isinstance(size, numbers.Number):
        size = np.array([size, size, size]).astype(np.float32)
    elif isinstance(size, list):
        size = np.array(size)
    elif isinstance(size, np.ndarray):
        size = size
    else:
        raise ValueError("size is either a number, or a list, or a np.ndarray")
    return size


def voxel_grid(pos, batch, size, start=None):
    """torch_geometric.nn.voxel_grid as sptr/utils.py:29 uses it: grid cell id of every point, the batch index acting as a
    4th coordinate of cell size 1 (torch_cluster grid: id = sum_d floor((p_d - start_d) / size_d) * prod(extent_<d))."""
    pos4 = torch.cat([pos, batch.to(pos.dtype)[:, None]], 1)
    size4 = torch.cat([torch.as_tensor(size, dtype=pos.dtype, device=pos.device).view(-1), torch.ones(1, dtype=pos.dtype, device=pos.device)])
    lo = pos4.min(0)[0]
    st = lo if start is None else torch.cat([torch.as_tensor(start, dtype=pos.dtype, device=pos.device).view(-1), lo[3:]])
    ext = torch.floor((pos4.max(0)[0] - st) / size4).long() + 1
    g = torch.floor((pos4 - st) / size4).long()
    mul = torch.cat([ext.new_ones(1), ext.cumprod(0)[:-1]])
    return (g * mul).sum(1)


def grid_sample(pos, batch, size, start, return_p2v=True, return_counts=True, return_unique=False):
    """sptr/utils.py:21-48."""
    cluster = voxel_grid(pos, batch, size, start=start)
    if return_p2v is False and return_counts is False:
        unique, cluster = torch.unique(cluster, sorted=True, return_inverse=True)
        return cluster
    unique, cluster, counts = torch.unique(cluster, sorted=True, return_inverse=True, return_counts=True)
    if return_p2v is False and return_counts is True:
        return cluster, counts.max().item(), counts
    n = unique.shape[0]
    k = counts.max().item()
    p2v_map = cluster.new_zeros(n, k)
    mask = torch.arange(k, device=cluster.device).unsqueeze(0) < counts.unsqueeze(-1)
    p2v_map[mask] = torch.argsort(cluster)
    if return_unique:
        return cluster, p2v_map, counts, unique
    return cluster, p2v_map, counts


def get_indices_params(xyz, batch, window_size, shift_win: bool):
    """sptr/utils.py:50-79: window partition of the points -> (index_0, index_0_offsets, n_max, index_1, index_1_offsets,
    sort_idx).  index_0_offsets carries the window boundaries for the fused operator (`_u2_windows`)."""
    if isinstance(window_size, (list, np.ndarray)):
        window_size = torch.from_numpy(np.asarray(window_size)).type_as(xyz).to(xyz.device)
    else:
        window_size = torch.tensor([window_size] * 3).type_as(xyz).to(xyz.device)
    if shift_win:
        v2p_map, k, counts = grid_sample(xyz + 1 / 2 * window_size, batch, window_size, start=xyz.min(0)[0], return_p2v=False,
                                         return_counts=True)
    else:
        v2p_map, k, counts = grid_sample(xyz, batch, window_size, start=None, return_p2v=False, return_counts=True)
    v2p_map, sort_idx = v2p_map.sort()
    n = counts.shape[0]
    N = v2p_map.shape[0]
    n_max = k
    index_0_offsets, index_1_offsets, index_0, index_1 = precompute_all(N, n, n_max, counts)
    windows = index_0_offsets._u2_windows
    index_0 = index_0.long()
    index_1 = index_1.long()
    index_0_offsets._u2_windows = windows
    return index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx


def scatter_softmax_csr(src: torch.Tensor, indptr: torch.Tensor, dim: int = -1):
    """sptr/utils.py:81-95: softmax of src [M, C] over the row segments indptr[i] .. indptr[i+1]-1 (torch_scatter's
    segment_csr / gather_csr replaced by torch segment arithmetic on the device)."""
    indptr = indptr.long()
    n_seg = indptr.shape[0] - 1
    seg = torch.repeat_interleave(torch.arange(n_seg, device=src.device), indptr[1:] - indptr[:-1])
    idx = seg[:, None].expand_as(src)
    mx = torch.full((n_seg, src.shape[1]), -float("inf"), dtype=src.dtype, device=src.device).scatter_reduce(0, idx, src, "amax")
    ex = (src - mx[seg]).exp()
    sm = torch.zeros((n_seg, src.shape[1]), dtype=src.dtype, device=src.device).index_add(0, seg, ex)
    return ex / sm[seg]
