"""CPU oracle of the SparseTransformer (`sptr`) window attention — TEST INFRASTRUCTURE, never imported by the product.

Restates, in plain torch index arithmetic (any dtype, fp64 for gradient checks; autograd gives the backward), what
third_party/SparseTransformer computes with its CUDA kernels.  PINNED on the reference: tests/test_gpu_sptr_ref.py compares
every function here with the reference's own kernels (oracle/_ref/libsptr_ref.so, compiled unmodified from /root/reference by
oracle/Makefile).  CPU-side pins: the reference's own test for precompute_all holds a
small known-answer case (third_party/SparseTransformer/test/test_precompute_all.py:9-19, 31-45, 67-70: counts [3, 2, 6]),
checked in tests/test_sptr_cpu.py; the attention operators have no stored vectors in the reference (its tests compare two
CUDA libraries on random data, test/test_attention_op_step1.py, test_relative_pos_encoding_op_step*.py), so for those the
oracle follows the kernels' indexing line by line and is additionally checked against an independent dense per-window
softmax attention.

Each function cites the reference lines it follows (paths relative to third_party/SparseTransformer/).
"""
import numpy as np
import torch


def precompute_all(N, n, n_max, counts):
    """sptr/functional.py:146-170 + src/sptr/precompute/precompute_cuda_kernel.cu:4-22.
    counts int [n] (points per window, points sorted by window) -> index_0_offsets [N+1], index_1_offsets [N], index_0 [M],
    index_1 [M]: pair m = sq_off[w] + i * len + t  <->  (query start + i, key start + t)."""
    counts = counts.long()
    offsets = torch.cat([counts.new_zeros(1), counts.cumsum(-1)])
    sq_offsets = torch.cat([counts.new_zeros(1), (counts ** 2).cumsum(-1)])
    M = int(sq_offsets[-1])
    index_0_offsets = torch.zeros(N, dtype=torch.int32)
    index_1_offsets = torch.zeros(N, dtype=torch.int32)
    index_0 = torch.zeros(M, dtype=torch.int32)
    index_1 = torch.zeros(M, dtype=torch.int32)
    for w in range(n):
        start, sv, length = int(offsets[w]), int(sq_offsets[w]), int(counts[w])
        for t in range(length):
            index_0_offsets[start + t] = sv + length * t
            index_1_offsets[start + t] = sv + t
            for i in range(length):
                index_0[sv + i * length + t] = start + i
                index_1[sv + i * length + t] = start + t
    index_0_offsets = torch.cat([index_0_offsets, torch.tensor([M], dtype=torch.int32)])
    return index_0_offsets, index_1_offsets, index_0, index_1


def precompute_all_fast(counts):
    """Vectorised form of the same layout (used for large random cases); checked against precompute_all."""
    counts = counts.long()
    n = counts.shape[0]
    offsets = torch.cat([counts.new_zeros(1), counts.cumsum(-1)])
    sq_offsets = torch.cat([counts.new_zeros(1), (counts ** 2).cumsum(-1)])
    N, M = int(offsets[-1]), int(sq_offsets[-1])
    win_of_pair = torch.repeat_interleave(torch.arange(n), counts ** 2)
    e = torch.arange(M) - sq_offsets[win_of_pair]
    length = counts[win_of_pair]
    index_0 = (offsets[win_of_pair] + e // length).int()
    index_1 = (offsets[win_of_pair] + e % length).int()
    win_of_pt = torch.repeat_interleave(torch.arange(n), counts)
    t = torch.arange(N) - offsets[win_of_pt]
    index_0_offsets = torch.cat([(sq_offsets[win_of_pt] + counts[win_of_pt] * t).int(), torch.tensor([M], dtype=torch.int32)])
    index_1_offsets = (sq_offsets[win_of_pt] + t).int()
    return index_0_offsets, index_1_offsets, index_0, index_1


def attention_step1(q, k, index_0, index_1):
    """src/sptr/attention/attention_cuda_kernel.cu:4-19: attn[m, h] = q[index_0[m], h, :] . k[index_1[m], h, :]."""
    return (q[index_0.long()] * k[index_1.long()]).sum(-1)


def dot_prod_with_idx(q, index_q, k, index_k, table_q, table_k, rel_idx):
    """src/sptr/rpe/relative_pos_encoding_cuda_kernel.cu:4-27: the relative-position part of the scores,
    q_i . (Tq[r0,0] + Tq[r1,1] + Tq[r2,2]) + k_j . (Tk[r0,0] + Tk[r1,1] + Tk[r2,2]); tables [L, 3, h, d], rel_idx [M, 3]."""
    r = rel_idx.long()
    tq = table_q[r[:, 0], 0] + table_q[r[:, 1], 1] + table_q[r[:, 2], 2]   # [M, h, d]
    tk = table_k[r[:, 0], 0] + table_k[r[:, 1], 1] + table_k[r[:, 2], 2]
    return (q[index_q.long()] * tq).sum(-1) + (k[index_k.long()] * tk).sum(-1)


def dot_prod_with_idx_all(q, index_q, k, index_k, table_q, table_k, rel_idx):
    """src/sptr/rpe/relative_pos_encoding_cuda_kernel.cu:116-145: content scores + relative-position scores."""
    return attention_step1(q, k, index_q, index_k) + dot_prod_with_idx(q, index_q, k, index_k, table_q, table_k, rel_idx)


def scatter_softmax_csr(src, indptr):
    """sptr/utils.py:81-95: softmax over the rows indptr[i] .. indptr[i+1]-1 of src [M, h], per column."""
    indptr = indptr.long()
    seg = torch.repeat_interleave(torch.arange(indptr.shape[0] - 1), indptr[1:] - indptr[:-1])
    n_seg = indptr.shape[0] - 1
    mx = torch.full((n_seg, src.shape[1]), -float("inf"), dtype=src.dtype).scatter_reduce(0, seg[:, None].expand_as(src), src, "amax")
    ex = (src - mx[seg]).exp()
    sm = torch.zeros((n_seg, src.shape[1]), dtype=src.dtype).index_add(0, seg, ex)
    return ex / sm[seg]


def attention_step2(attn, v, index_0, index_1, N):
    """src/sptr/attention/attention_cuda_kernel.cu:77-99: out[i, h, :] = sum over the pairs of query i of attn[m, h] v[index_1[m], h, :]."""
    out = torch.zeros((N,) + tuple(v.shape[1:]), dtype=v.dtype)
    return out.index_add(0, index_0.long(), attn[:, :, None] * v[index_1.long()])


def attention_step2_with_rel_pos_value(attn, v, index_0, index_1, table, rel_idx, N):
    """src/sptr/rpe/relative_pos_encoding_cuda_kernel.cu (attention_step2_with_rel_pos_value_forward): values get the
    relative-position rows added: out[i] = sum_m attn[m] (v[j] + Tv[r0,0] + Tv[r1,1] + Tv[r2,2])."""
    r = rel_idx.long()
    tv = table[r[:, 0], 0] + table[r[:, 1], 1] + table[r[:, 2], 2]
    out = torch.zeros((N,) + tuple(v.shape[1:]), dtype=v.dtype)
    return out.index_add(0, index_0.long(), attn[:, :, None] * (v[index_1.long()] + tv))


def window_attention(q, k, v, counts, rel_idx=None, table_q=None, table_k=None, table_v=None):
    """The chain sparse_self_attention runs on window-sorted points (sptr/modules.py:36-62 without the sort / un-sort)."""
    N = q.shape[0]
    i0o, i1o, i0, i1 = precompute_all_fast(counts)
    if rel_idx is not None:
        s = dot_prod_with_idx_all(q, i0, k, i1, table_q, table_k, rel_idx)
    else:
        s = attention_step1(q, k, i0, i1)
    p = scatter_softmax_csr(s, i0o)
    if rel_idx is not None:
        return attention_step2_with_rel_pos_value(p, v, i0, i1, table_v, rel_idx, N)
    return attention_step2(p, v, i0, i1, N)


def dense_window_attention(q, k, v, counts, rel_idx=None, table_q=None, table_k=None, table_v=None):
    """Independent formulation for checking the oracle itself: one dense softmax attention per window."""
    outs, start, sq = [], 0, 0
    for n in counts.tolist():
        qs, ks, vs = q[start:start + n], k[start:start + n], v[start:start + n]          # [n, h, d]
        s = torch.einsum("ihd,jhd->ijh", qs, ks)
        if rel_idx is not None:
            r = rel_idx[sq:sq + n * n].long().view(n, n, 3)
            tq = table_q[r[..., 0], 0] + table_q[r[..., 1], 1] + table_q[r[..., 2], 2]   # [n, n, h, d]
            tk = table_k[r[..., 0], 0] + table_k[r[..., 1], 1] + table_k[r[..., 2], 2]
            tv = table_v[r[..., 0], 0] + table_v[r[..., 1], 1] + table_v[r[..., 2], 2]
            s = s + torch.einsum("ihd,ijhd->ijh", qs, tq) + torch.einsum("jhd,ijhd->ijh", ks, tk)
            p = torch.softmax(s, dim=1)
            outs.append(torch.einsum("ijh,ijhd->ihd", p, vs[None] + tv))
        else:
            p = torch.softmax(s, dim=1)
            outs.append(torch.einsum("ijh,jhd->ihd", p, vs))
        start += n
        sq += n * n
    return torch.cat(outs, 0)


def voxel_grid_cluster(pos, batch, size, start=None):
    """torch_geometric.nn.voxel_grid as sptr/utils.py:29 calls it (torch_cluster grid: the batch index is appended as a 4th
    coordinate with voxel size 1; cluster = sum_d floor((p_d - start_d) / size_d) * prod of the extents before d).  Only the
    PARTITION matters downstream (sptr/utils.py:33-36 takes unique + inverse)."""
    pos4 = torch.cat([pos, batch.to(pos.dtype)[:, None]], 1)
    size4 = torch.cat([torch.as_tensor(size, dtype=pos.dtype), torch.ones(1, dtype=pos.dtype)])
    st = pos4.min(0)[0] if start is None else torch.cat([torch.as_tensor(start, dtype=pos.dtype), pos4[:, 3:].min(0)[0]])
    end = pos4.max(0)[0]
    g = torch.floor((pos4 - st) / size4).long()
    ext = torch.floor((end - st) / size4).long() + 1
    cluster, mul = torch.zeros(pos.shape[0], dtype=torch.long), 1
    for d in range(4):
        cluster = cluster + g[:, d] * mul
        mul = mul * int(ext[d])
    return cluster


def get_indices_params(xyz, batch, window_size, shift_win):
    """sptr/utils.py:50-79: window partition -> (index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx)."""
    ws = torch.as_tensor(np.asarray(window_size, dtype=np.float32) if not np.isscalar(window_size) else np.full(3, window_size, np.float32)).to(xyz.dtype)
    if shift_win:
        cluster = voxel_grid_cluster(xyz + 0.5 * ws, batch, ws, start=xyz.min(0)[0])
    else:
        cluster = voxel_grid_cluster(xyz, batch, ws, start=None)
    _, v2p, counts = torch.unique(cluster, sorted=True, return_inverse=True, return_counts=True)
    v2p_sorted, sort_idx = torch.sort(v2p, stable=True)
    n_max = int(counts.max())
    i0o, i1o, i0, i1 = precompute_all_fast(counts)
    return i0.long(), i0o, n_max, i1.long(), i1o, sort_idx, counts


# ------------------------------------------------------------------ module-level facade (what the SphereFormer blocks import)
def to_3d_numpy(size):
    """sptr/utils.py:9-19."""
    import numbers
    if isinstance(size, numbers.Number):
        return np.array([size, size, size]).astype(np.float32)
    if isinstance(size, list):
        return np.array(size)
    if isinstance(size, np.ndarray):
        return size
    raise ValueError("size is either a number, or a list, or a np.ndarray")


class SparseTrTensor(object):
    """sptr/__init__.py:4-33."""

    def __init__(self, query_feats, query_indices, spatial_shape, batch_size, key_feats=None, value_feats=None, key_indices=None):
        self.query_feats, self.key_feats, self.value_feats = query_feats, key_feats, value_feats
        self.query_indices, self.key_indices = query_indices, key_indices
        self.spatial_shape, self.batch_size = spatial_shape, batch_size
        self.indice_dict = {}

    def find_indice_params(self, key):
        return None if key is None else self.indice_dict.get(key)


def get_indices_params_ref(xyz, batch, window_size, shift_win):
    """The reference's 6-tuple (sptr/utils.py:79)."""
    i0, i0o, n_max, i1, i1o, sort_idx, _ = get_indices_params(xyz, batch, window_size, shift_win)
    return i0, i0o, n_max, i1, i1o, sort_idx


def sparse_self_attention(query, key, value, xyz, index_0, index_0_offsets, n_max, index_1, index_1_offsets, sort_idx,
                          window_size, shift_win, pe_type='none', rel_query=False, rel_key=False, rel_value=False,
                          quant_size=None, quant_grid_length=None, relative_pos_query_table=None,
                          relative_pos_key_table=None, relative_pos_value_table=None, split_func=None):
    """sptr/modules.py:11-62 line by line over the oracle operators (CPU, any float dtype)."""
    query, key, value, xyz_ctg = query[sort_idx], key[sort_idx], value[sort_idx], xyz[sort_idx]
    N = query.shape[0]
    if pe_type == 'contextual' and rel_query and rel_key:
        ws = torch.from_numpy(np.asarray(window_size)).to(xyz.dtype)
        shift_size = 1 / 2 * ws if shift_win else 0.0
        xyz_quant = (xyz_ctg - xyz_ctg.min(0)[0] + shift_size) % ws
        xyz_quant = torch.div(xyz_quant, torch.from_numpy(np.asarray(quant_size)).to(xyz.dtype), rounding_mode='floor')
        relative_position = xyz_quant[index_0.long()] - xyz_quant[index_1.long()]
        relative_position_index = relative_position + quant_grid_length - 1
        if split_func:
            relative_position_index = split_func(xyz_ctg, index_0, index_1, relative_position_index.clone())
            relative_position_index = torch.clamp(relative_position_index, 0, 2 * quant_grid_length - 1)
        relative_position_index = relative_position_index.int()
        attn_flat = dot_prod_with_idx_all(query, index_0, key, index_1, relative_pos_query_table, relative_pos_key_table,
                                          relative_position_index)
    else:
        attn_flat = attention_step1(query, key, index_0, index_1)
    softmax_attn_flat = scatter_softmax_csr(attn_flat, index_0_offsets)
    if pe_type == 'contextual' and rel_value:
        x = attention_step2_with_rel_pos_value(softmax_attn_flat, value, index_0, index_1, relative_pos_value_table,
                                               relative_position_index, N)
    else:
        x = attention_step2(softmax_attn_flat, value, index_0, index_1, N)
    out = torch.empty_like(x)
    out[sort_idx] = x
    return out


def as_sptr_module():
    from types import SimpleNamespace
    return SimpleNamespace(to_3d_numpy=to_3d_numpy, SparseTrTensor=SparseTrTensor, sparse_self_attention=sparse_self_attention,
                           get_indices_params=get_indices_params_ref)
