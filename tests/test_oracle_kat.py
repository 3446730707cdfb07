"""Pins for the CPU oracle (oracle/ts_oracle.py).  The reference ships no golden vectors for
this path ("parity unpinned"), so the oracle is pinned by known-answer tests:
dense-equivalence identities, brute-force dictionary kernel maps,
hand-computed micro cases and fp64 gradcheck."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF


def _full_grid(B, X, Y, Z, rng):
    g = np.stack(np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).reshape(-1, 3)
    c = np.concatenate([np.concatenate([g, np.full((g.shape[0], 1), b)], 1) for b in range(B)]).astype(np.int32)
    rng.shuffle(c)
    return torch.from_numpy(c)


def _to_sparse_feats(dense, coords):
    c = coords.long()
    return dense[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]].contiguous()


def _to_dense(feats, coords, shape):
    B, C, X, Y, Z = shape
    out = torch.zeros(shape, dtype=feats.dtype)
    c = coords.long()
    out[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = feats
    return out


def test_hash_known_answer(oracle):
    """FNV-1a over 4 words folded to 60 bits, recomputed here in pure Python."""
    def fnv(c):
        h = 14695981039346656037
        for v in c:
            h ^= (v & 0xFFFFFFFF)
            h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return (h >> 60) ^ (h & 0xFFFFFFFFFFFFFFF)
    c = torch.tensor([[0, 0, 0, 0], [1, 2, 3, 0], [-1, 5, 7, 1], [2047, 4095, 63, 3]], dtype=torch.int)
    assert oracle.sphash(c).tolist() == [fnv(r) for r in c.tolist()]
    off = torch.tensor([[1, 0, -1], [0, 0, 0]], dtype=torch.int)
    got = oracle.sphash(c, off)
    for k, o in enumerate(off.tolist()):
        assert got[k].tolist() == [fnv([r[0] + o[0], r[1] + o[1], r[2] + o[2], r[3]]) for r in c.tolist()]


def test_kernel_offsets_order(oracle):
    o3 = oracle.get_kernel_offsets(3).tolist()
    assert o3[0] == [-1, -1, -1] and o3[1] == [0, -1, -1] and o3[13] == [0, 0, 0] and o3[26] == [1, 1, 1]  # x fastest
    o2 = oracle.get_kernel_offsets(2, 4).tolist()
    assert o2 == [[0, 0, 0], [0, 0, 4], [0, 4, 0], [0, 4, 4], [4, 0, 0], [4, 0, 4], [4, 4, 0], [4, 4, 4]]  # z fastest


def test_hashquery_count_micro(oracle):
    ref = torch.tensor([50, 10, 30, 10], dtype=torch.long)
    q = torch.tensor([[10, 99], [30, 50]], dtype=torch.long)
    assert oracle.sphashquery(q, ref).tolist() == [[1, -1], [2, 0]]  # first duplicate wins
    idx = torch.tensor([0, 2, 2, -1, 1, 2], dtype=torch.int)
    assert oracle.spcount(idx, 4).tolist() == [1, 1, 3, 0]


def test_voxelize_micro(oracle):
    feats = torch.tensor([[1., 2.], [3., 4.], [5., 6.], [7., 8.]])
    idx = torch.tensor([1, 0, 1, -1], dtype=torch.int)
    counts = oracle.spcount(idx, 2)
    out = oracle.spvoxelize(feats, idx, counts)
    assert torch.allclose(out, torch.tensor([[3., 4.], [3., 4.]]))


@pytest.mark.parametrize("thin", [1.0, 0.3])
def test_submanifold_conv_equals_dense_conv3d(oracle, thin):
    rng = np.random.default_rng(0)
    B, X, Y, Z, Ci, Co = 2, 6, 4, 8, 3, 5
    coords = _full_grid(B, X, Y, Z, rng)
    coords = coords[torch.from_numpy(rng.random(coords.shape[0]) < thin)]
    dense = torch.zeros(B, Ci, X, Y, Z, dtype=torch.float64)
    feats = torch.from_numpy(rng.standard_normal((coords.shape[0], Ci)))
    dense = _to_dense(feats, coords, dense.shape)
    conv = oracle.Conv3d(Ci, Co, 3).double()
    y = conv(oracle.SparseTensor(feats, coords))
    Wd = conv.kernel.detach().view(3, 3, 3, Ci, Co).permute(4, 3, 2, 1, 0)
    yd = TF.conv3d(dense, Wd, padding=1)
    assert torch.allclose(y.F, _to_sparse_feats(yd, coords), atol=1e-12)


@pytest.mark.parametrize("thin", [1.0, 0.3])
def test_strided_and_transposed_conv_equal_dense(oracle, thin):
    rng = np.random.default_rng(1)
    B, X, Y, Z, Ci, Co = 2, 6, 4, 8, 3, 5
    coords = _full_grid(B, X, Y, Z, rng)
    coords = coords[torch.from_numpy(rng.random(coords.shape[0]) < thin)]
    feats = torch.from_numpy(rng.standard_normal((coords.shape[0], Ci)))
    dense = _to_dense(feats, coords, (B, Ci, X, Y, Z))
    down = oracle.Conv3d(Ci, Co, 2, 2).double()
    x = oracle.SparseTensor(feats, coords)
    x.cmaps[x.s] = x.C
    y = down(x)
    assert y.s == (2, 2, 2)
    want_coords = torch.unique(torch.cat([coords[:, :3] // 2 * 2, coords[:, 3:]], 1)[:, [3, 0, 1, 2]], dim=0)[:, [1, 2, 3, 0]]
    assert torch.equal(y.C, want_coords)
    Wd2 = down.kernel.detach().view(2, 2, 2, Ci, Co).permute(4, 3, 0, 1, 2)
    yd = TF.conv3d(dense, Wd2, stride=2)
    c2 = y.C.clone()
    c2[:, :3] //= 2
    assert torch.allclose(y.F, _to_sparse_feats(yd, c2), atol=1e-12)
    # transposed conv back to stride 1: rows/coords are those of the encoder tensor
    up = oracle.Conv3d(Co, Ci, 2, 2, transposed=True).double()
    z = up(y)
    assert z.s == (1, 1, 1) and z.C is x.C
    Wt = up.kernel.detach().view(2, 2, 2, Co, Ci).permute(3, 4, 0, 1, 2)
    y_dense = _to_dense(y.F.detach(), c2, (B, Co, X // 2, Y // 2, Z // 2))
    zd = TF.conv_transpose3d(y_dense, Wt, stride=2)
    assert torch.allclose(z.F, _to_sparse_feats(zd, coords), atol=1e-12)


def test_devoxelize_equals_grid_sample(oracle):
    rng = np.random.default_rng(2)
    B, X, Y, Z, C = 1, 5, 6, 7, 4
    coords = _full_grid(B, X, Y, Z, rng)
    dense = torch.from_numpy(rng.standard_normal((B, C, X, Y, Z)))
    feats = _to_sparse_feats(dense, coords)
    n = 200
    p = torch.from_numpy(rng.random((n, 3)) * (np.array([X, Y, Z]) - 1.001))
    pts = torch.cat([p, torch.zeros(n, 1, dtype=torch.float64)], 1)
    key = torch.cat([torch.floor(pts[:, :3]).int(), pts[:, -1].int().view(-1, 1)], 1)
    idx = oracle.sphashquery(oracle.sphash(key, oracle.get_kernel_offsets(2, 1, 1)), oracle.sphash(coords))
    w = oracle.calc_ti_weights(pts, idx, scale=1).t().contiguous()
    out = oracle.spdevoxelize(feats, idx.t().contiguous(), w)
    grid = torch.stack([2 * p[:, 2] / (Z - 1) - 1, 2 * p[:, 1] / (Y - 1) - 1, 2 * p[:, 0] / (X - 1) - 1], -1)
    want = TF.grid_sample(dense, grid.view(1, n, 1, 1, 3), mode="bilinear", align_corners=True).view(C, n).t()
    assert torch.allclose(out, want, atol=1e-6)


def test_point_to_voxel_equals_avg_pool(oracle):
    from u2mkd_b200 import models
    fam = models.build_family(oracle.as_torchsparse_modules()["torchsparse"])
    rng = np.random.default_rng(3)
    B, X, Y, Z, C = 2, 4, 6, 8, 3
    coords = _full_grid(B, X, Y, Z, rng)
    dense = torch.from_numpy(rng.standard_normal((B, C, X, Y, Z)).astype(np.float32))
    z = oracle.PointTensor(_to_sparse_feats(dense, coords), coords.float())
    c2 = torch.unique(torch.cat([coords[:, :3] // 2 * 2, coords[:, 3:]], 1), dim=0)
    x = oracle.SparseTensor(torch.zeros(c2.shape[0], 1), c2, 2)
    out = fam.point_to_voxel(x, z)
    assert z.additional_features["counts"][(2, 2, 2)].eq(8).all()
    pooled = TF.avg_pool3d(dense, 2)
    cc = c2.clone()
    cc[:, :3] //= 2
    assert torch.allclose(out.F, _to_sparse_feats(pooled, cc), atol=1e-6)


def test_kernel_map_equals_bruteforce_dict(oracle):
    rng = np.random.default_rng(4)
    c = np.unique(np.concatenate([rng.integers(0, 12, size=(900, 3)), rng.integers(0, 2, size=(900, 1))], 1), axis=0)
    c = torch.from_numpy(c.astype(np.int32))
    kmap, out_c = oracle.build_kernel_map(c, (1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    table = {tuple(r): i for i, r in enumerate(c.tolist())}
    want, sizes = [], []
    for off in oracle.get_kernel_offsets(3).tolist():
        rows = [(table[(x + off[0], y + off[1], z + off[2], b)], o) for o, (x, y, z, b) in enumerate(c.tolist())
                if (x + off[0], y + off[1], z + off[2], b) in table]
        want += rows
        sizes.append(len(rows))
    assert kmap[0].tolist() == [list(r) for r in want]
    assert kmap[1].tolist() == sizes and kmap[2] == (c.shape[0], c.shape[0])


def test_gradcheck_fp64(oracle):
    rng = np.random.default_rng(5)
    c = np.unique(np.concatenate([rng.integers(0, 5, size=(60, 3)), np.zeros((60, 1), int)], 1), axis=0).astype(np.int32)
    c = torch.from_numpy(c)
    kmap, _ = oracle.build_kernel_map(c, (1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    f = torch.from_numpy(rng.standard_normal((c.shape[0], 2))).requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((27, 2, 3))).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, b: oracle._ConvolutionFn.apply(a, b, kmap[0], kmap[1], kmap[2], False), (f, w))
    idx = torch.from_numpy(rng.integers(-1, 7, size=c.shape[0]).astype(np.int32))
    counts = oracle.spcount(idx, 7)
    assert torch.autograd.gradcheck(lambda a: oracle.spvoxelize(a, idx, counts), (f,))
    idx8 = torch.from_numpy(rng.integers(-1, c.shape[0], size=(20, 8)).astype(np.int32))
    w8 = torch.from_numpy(rng.random((20, 8)))
    assert torch.autograd.gradcheck(lambda a: oracle.spdevoxelize(a, idx8, w8), (f,))


def test_sparse_quantize_and_collate(oracle):
    pts = np.array([[0.1, 0.1, 0.1], [0.9, 0.2, 0.3], [1.5, 0.0, 0.0], [0.2, 0.2, 0.2], [1.9, 0.9, 0.9]])
    coords, ind, inv = oracle.sparse_quantize(pts, 1.0, return_index=True, return_inverse=True)
    assert coords.tolist() == [[0, 0, 0], [1, 0, 0]] and ind.tolist() == [0, 2] and inv.tolist() == [0, 0, 1, 0, 1]
    a = oracle.SparseTensor(torch.ones(2, 3), torch.zeros(2, 3, dtype=torch.int))
    b = oracle.SparseTensor(torch.ones(1, 3), torch.ones(1, 3, dtype=torch.int))
    out = oracle.sparse_collate([a, b])
    assert out.C.tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 1, 1]] and out.F.shape == (3, 3)
