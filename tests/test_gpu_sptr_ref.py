"""f1 parity against the REFERENCE'S OWN CUDA KERNELS: oracle/_ref/libsptr_ref.so is third_party/SparseTransformer/src/sptr/
{attention,rpe,precompute}/*_cuda_kernel.cu compiled unmodified from /root/reference (oracle/Makefile, `sptr_ref`; built by
__graft_entry__.build() wherever the reference tree exists, and shipped to the GPU box with the snapshot).  Checks, at the
shapes of the reference's operator tests (N ~ 3500 points, 150 windows, 6 heads of 16, L = 31):
  * precompute_all: product == reference kernel, bit for bit;
  * every reference kernel (scores, value step, both backward pairs) == oracle/sptr_oracle.py (fp64) to fp32 accuracy —
    this PINS the oracle on reference outputs;
  * the product's single fused forward / backward kernel == the reference's three-kernel chain end to end (reference scores
    -> segment softmax -> reference value step; backward through the reference's backward kernels), all six gradients.
Skipped when the library has not been built (no reference tree at build time)."""
import numpy as np
import pytest
import torch

from oracle import sptr_oracle as so
from oracle import sptr_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not sptr_ref.available(), reason="oracle/_ref/libsptr_ref.so not built")]


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def case():
    rng = np.random.default_rng(1)
    n_win, h, d, L = 150, 6, 16, 31
    counts = torch.from_numpy(rng.integers(1, 48, size=n_win).astype(np.int64))
    N, M = int(counts.sum()), int((counts ** 2).sum())
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.rand(N, h, d, generator=g) for _ in range(3))            # torch.rand as in the reference's tests
    tq, tk, tv = (torch.rand(L, 3, h, d, generator=g) for _ in range(3))
    rel = torch.from_numpy(rng.integers(0, L, size=(M, 3)).astype(np.int32))
    i0o, i1o, i0, i1 = so.precompute_all_fast(counts)
    return dict(counts=counts, N=N, M=M, h=h, d=d, L=L, n_max=int(counts.max()), q=q * d ** -0.5, k=k, v=v, tq=tq, tk=tk, tv=tv,
                rel=rel, i0o=i0o, i1o=i1o, i0=i0, i1=i1)


def test_precompute_all_product_equals_reference_kernel(cuda_lib, case):
    from u2mkd_b200 import sptr
    c = case
    ref = sptr_ref.precompute_all(c["N"], c["counts"].shape[0], c["n_max"], c["counts"].cuda())
    got = sptr.precompute_all(c["N"], c["counts"].shape[0], c["n_max"], c["counts"].int().cuda())
    for a, b, w in zip(got, ref, (c["i0o"], c["i1o"], c["i0"], c["i1"])):
        assert torch.equal(a.cpu(), b.cpu()) and torch.equal(b.cpu(), w)
    # the reference's known-answer case (test/test_precompute_all.py): counts [3, 2, 6]
    kat = sptr_ref.precompute_all(11, 3, 6, torch.tensor([3, 2, 6], dtype=torch.int32).cuda())
    assert kat[0].tolist() == [0, 3, 6, 9, 11, 13, 19, 25, 31, 37, 43, 49] and kat[1].tolist() == [0, 1, 2, 9, 10, 13, 14, 15, 16, 17, 18]


def test_reference_kernels_pin_the_oracle(cuda_lib, case):
    c = case
    cu = {k_: (v_.cuda() if torch.is_tensor(v_) else v_) for k_, v_ in c.items()}
    d64 = {k_: (v_.double() if torch.is_tensor(v_) and v_.is_floating_point() else v_) for k_, v_ in c.items()}
    # scores
    s_ref = sptr_ref.dot_prod_with_idx_all_forward(cu["q"], cu["k"], cu["i0"], cu["i0o"], cu["i1"], cu["tq"], cu["tk"], cu["rel"], c["n_max"])
    leaves = [d64[n].clone().requires_grad_(True) for n in ("q", "k", "tq", "tk")]
    s_or = so.dot_prod_with_idx_all(leaves[0], c["i0"], leaves[1], c["i1"], leaves[2], leaves[3], c["rel"])
    assert rel_err(s_ref, s_or) < 1e-5
    g = torch.rand(c["M"], c["h"], generator=torch.Generator().manual_seed(2))
    s_or.backward(g.double())
    grads = sptr_ref.dot_prod_with_idx_all_backward(g.cuda(), cu["q"], cu["k"], cu["i0"], cu["i0o"], cu["i1"], cu["i1o"], cu["tq"], cu["tk"],
                                                    cu["rel"], c["n_max"])
    for got, leaf, name in zip(grads, leaves, ("dq", "dk", "dtable_q", "dtable_k")):
        assert rel_err(got, leaf.grad) < 1e-4, (name, rel_err(got, leaf.grad))
    # value step
    p = so.scatter_softmax_csr(s_or.detach(), c["i0o"])
    leaves2 = [p.clone().requires_grad_(True), d64["v"].clone().requires_grad_(True), d64["tv"].clone().requires_grad_(True)]
    o_or = so.attention_step2_with_rel_pos_value(leaves2[0], leaves2[1], c["i0"], c["i1"], leaves2[2], c["rel"], c["N"])
    o_ref = sptr_ref.attention_step2_with_rel_pos_value_forward(p.float().cuda(), cu["v"], cu["i0o"], cu["i1"], cu["tv"], cu["rel"], c["n_max"])
    assert rel_err(o_ref, o_or) < 1e-5
    go = torch.rand(c["N"], c["h"], c["d"], generator=torch.Generator().manual_seed(3))
    o_or.backward(go.double())
    g2 = sptr_ref.attention_step2_with_rel_pos_value_backward(go.cuda(), p.float().cuda(), cu["v"], cu["i0"], cu["i0o"], cu["i1"], cu["i1o"],
                                                              cu["tv"], cu["rel"], c["n_max"])
    for got, leaf, name in zip(g2, leaves2, ("dattn", "dv", "dtable_v")):
        assert rel_err(got, leaf.grad) < 1e-4, (name, rel_err(got, leaf.grad))


def test_fused_kernels_equal_the_reference_chain(cuda_lib, case):
    from u2mkd_b200.sptr import functional as F
    from u2mkd_b200.sptr.utils import scatter_softmax_csr
    c = case
    cu = {k_: (v_.cuda() if torch.is_tensor(v_) else v_) for k_, v_ in c.items()}
    # reference chain, forward
    s_ref = sptr_ref.dot_prod_with_idx_all_forward(cu["q"], cu["k"], cu["i0"], cu["i0o"], cu["i1"], cu["tq"], cu["tk"], cu["rel"], c["n_max"])
    s_leaf = s_ref.clone().requires_grad_(True)
    p_ref = scatter_softmax_csr(s_leaf, cu["i0o"].long(), dim=0)          # torch_scatter's role: segment softmax (torch ops)
    o_ref = sptr_ref.attention_step2_with_rel_pos_value_forward(p_ref.detach().contiguous(), cu["v"], cu["i0o"], cu["i1"], cu["tv"], cu["rel"], c["n_max"])
    # reference chain, backward
    go = torch.rand(c["N"], c["h"], c["d"], generator=torch.Generator().manual_seed(4)).cuda()
    d_attn, dv_ref, dtv_ref = sptr_ref.attention_step2_with_rel_pos_value_backward(go, p_ref.detach().contiguous(), cu["v"], cu["i0"], cu["i0o"], cu["i1"],
                                                                                   cu["i1o"], cu["tv"], cu["rel"], c["n_max"])
    p_ref.backward(d_attn)
    dq_ref, dk_ref, dtq_ref, dtk_ref = sptr_ref.dot_prod_with_idx_all_backward(s_leaf.grad.contiguous(), cu["q"], cu["k"], cu["i0"], cu["i0o"], cu["i1"],
                                                                              cu["i1o"], cu["tq"], cu["tk"], cu["rel"], c["n_max"])
    # product: one kernel each way
    leaves = [cu[n].clone().requires_grad_(True) for n in ("q", "k", "v", "tq", "tk", "tv")]
    win_off, sq_off = F.window_offsets(cu["counts"])
    o = F.window_attention(leaves[0], leaves[1], leaves[2], win_off, sq_off, c["counts"].shape[0], cu["rel"], leaves[3], leaves[4], leaves[5])
    o.backward(go)
    assert rel_err(o, o_ref) < 1e-5, rel_err(o, o_ref)
    for leaf, want, name in zip(leaves, (dq_ref, dk_ref, dv_ref, dtq_ref, dtk_ref, dtv_ref), ("dq", "dk", "dv", "dtable_q", "dtable_k", "dtable_v")):
        assert rel_err(leaf.grad, want) < 1e-4, (name, rel_err(leaf.grad, want))
