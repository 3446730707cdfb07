"""CPU checks of the sptr oracle (oracle/sptr_oracle.py): the reference's own known-answer case for precompute_all
(third_party/SparseTransformer/test/test_precompute_all.py:9-45: counts [3, 2, 6]) and the oracle's attention chain
against an independent dense per-window softmax attention."""
import numpy as np
import torch

from oracle import sptr_oracle as so


def test_precompute_all_reference_known_answer():
    """test_precompute_all.py: v2p_map [1,0,0,2,0,2,2,1,2,2,2] sorted, counts [3,2,6]; the expected arrays are the ones
    that test derives: index_0_offsets = [0] + cumsum(counts[v2p]), index_1_offsets = sq_off[v2p] + rank inside the
    window, and pairs (query-major inside a window)."""
    counts = torch.tensor([3, 2, 6], dtype=torch.int32)
    v2p = torch.tensor([1, 0, 0, 2, 0, 2, 2, 1, 2, 2, 2]).sort().values.long()
    N, n, k = 11, 3, 6
    mask = torch.arange(k)[None].expand(n, -1) < counts[:, None]
    to_add = torch.arange(k)[None].expand(n, -1)[mask]
    want_i1o = torch.cat([torch.zeros(1, dtype=torch.long), (counts.long() ** 2).cumsum(-1)])[v2p] + to_add
    want_i0o = torch.cat([torch.zeros(1, dtype=torch.long), counts.long()[v2p].cumsum(-1)])
    i0o, i1o, i0, i1 = so.precompute_all(N, n, k, counts)
    assert i0o.tolist() == want_i0o.tolist() == [0, 3, 6, 9, 11, 13, 19, 25, 31, 37, 43, 49]
    assert i1o.tolist() == want_i1o.tolist() == [0, 1, 2, 9, 10, 13, 14, 15, 16, 17, 18]
    # pointops.precompute_index_pairs of that test: for query i of a window, its keys in order
    assert i0[:9].tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2] and i1[:9].tolist() == [0, 1, 2, 0, 1, 2, 0, 1, 2]
    assert i0[9:13].tolist() == [3, 3, 4, 4] and i1[9:13].tolist() == [3, 4, 3, 4]
    assert i0[13:19].tolist() == [5] * 6 and i1[13:19].tolist() == [5, 6, 7, 8, 9, 10]
    f = so.precompute_all_fast(counts)
    for a, b in zip(f, (i0o, i1o, i0, i1)):
        assert torch.equal(a, b)


def test_fast_pairs_match_the_kernel_restatement_on_random_windows():
    rng = np.random.default_rng(0)
    counts = torch.from_numpy(rng.integers(1, 9, size=40).astype(np.int32))
    N = int(counts.sum())
    slow = so.precompute_all(N, 40, int(counts.max()), counts)
    fast = so.precompute_all_fast(counts)
    for a, b in zip(slow, fast):
        assert torch.equal(a, b)


def test_oracle_chain_equals_dense_window_attention():
    torch.manual_seed(0)
    rng = np.random.default_rng(1)
    counts = torch.from_numpy(rng.integers(1, 12, size=15).astype(np.int64))
    N, M, h, d, L = int(counts.sum()), int((counts ** 2).sum()), 3, 16, 7
    q, k, v = (torch.randn(N, h, d, dtype=torch.float64) for _ in range(3))
    tq, tk, tv = (torch.randn(L, 3, h, d, dtype=torch.float64) * 0.3 for _ in range(3))
    rel = torch.from_numpy(rng.integers(0, L, size=(M, 3)).astype(np.int32))
    for args in ((), (rel, tq, tk, tv)):
        a = so.window_attention(q, k, v, counts, *args)
        b = so.dense_window_attention(q, k, v, counts, *args)
        assert float((a - b).abs().max()) < 1e-12
    # softmax rows sum to one
    i0o = so.precompute_all_fast(counts)[0]
    p = so.scatter_softmax_csr(torch.randn(M, h, dtype=torch.float64), i0o)
    seg = torch.repeat_interleave(torch.arange(N), (i0o[1:] - i0o[:-1]).long())
    assert float((torch.zeros(N, h, dtype=torch.float64).index_add(0, seg, p) - 1).abs().max()) < 1e-12


def test_window_partition_is_a_partition_by_grid_cell():
    rng = np.random.default_rng(2)
    xyz = torch.from_numpy(rng.uniform(0, 10, size=(500, 3)).astype(np.float32))
    batch = torch.from_numpy(rng.integers(0, 2, size=500))
    for shift in (False, True):
        i0, i0o, n_max, i1, i1o, sort_idx, counts = so.get_indices_params(xyz, batch, np.array([2.5, 2.5, 2.5], np.float32), shift)
        assert int(counts.sum()) == 500 and n_max == int(counts.max()) and sorted(sort_idx.tolist()) == list(range(500))
        ws = 2.5
        base = xyz.min(0)[0]
        cell = torch.floor((xyz + (0.5 * ws if shift else 0.0) - base) / ws).long()
        key = [tuple(c) + (int(b),) for c, b in zip(cell.tolist(), batch.tolist())]
        start = 0
        for n in counts.tolist():      # every window = all points of exactly one (cell, batch)
            members = sort_idx[start:start + n].tolist()
            assert len({key[m] for m in members}) == 1
            assert sum(1 for kk in key if kk == key[members[0]]) == n
            start += n
