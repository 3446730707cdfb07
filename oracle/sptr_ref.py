"""ctypes access to oracle/_ref/libsptr_ref.so — the REFERENCE's own sptr CUDA kernels (third_party/SparseTransformer/src/sptr/
{attention,rpe,precompute}/*_cuda_kernel.cu), compiled from where they lie by oracle/Makefile (`make -C oracle sptr_ref`).
TEST INFRASTRUCTURE: used by tests/test_gpu_sptr_ref.py to pin oracle/sptr_oracle.py and the product's fused kernels on the
real reference arithmetic.  Every wrapper reproduces the tensor layouts the reference's Python side prepares before the call
(sptr/functional.py: the forward kernels take q / k / tables / rel_idx TRANSPOSED, file:line cited per function)."""
import ctypes
import os

import torch

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsptr_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(SO)
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous(), (t.device, t.is_contiguous())
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    return t.int().contiguous()


def precompute_all(N, n, n_max, counts):
    """sptr/functional.py:146-170."""
    counts = counts.int().contiguous()
    offsets = torch.cat([counts.new_zeros(1), counts.cumsum(-1)], 0).int().contiguous()
    sq_offsets = torch.cat([counts.new_zeros(1), (counts.long() ** 2).cumsum(-1)], 0).int().contiguous()
    M = int(sq_offsets[-1])
    dev = counts.device
    i0o, i1o = torch.zeros(N, dtype=torch.int32, device=dev), torch.zeros(N, dtype=torch.int32, device=dev)
    i0, i1 = torch.zeros(M, dtype=torch.int32, device=dev), torch.zeros(M, dtype=torch.int32, device=dev)
    lib().precompute_all_cuda_launcher(int(N), int(n), ctypes.c_uint(int(n_max)), _p(counts), _p(offsets), _p(sq_offsets), _p(i0o), _p(i1o), _p(i0), _p(i1))
    torch.cuda.synchronize()
    return torch.cat([i0o, torch.tensor([M], dtype=torch.int32, device=dev)]), i1o, i0, i1


def dot_prod_with_idx_all_forward(q, k, index_q, index_q_offsets, index_k, table_q, table_k, rel_idx, n_max):
    """sptr/functional.py:253-292: scores [M, h] = content + relative-position terms."""
    N, h, d = q.shape
    M, L = index_k.shape[0], table_q.shape[0]
    out = torch.zeros(h, M, dtype=torch.float32, device=q.device)
    qt, kt = q.permute(1, 2, 0).contiguous(), k.permute(1, 2, 0).contiguous()
    tq, tk = table_q.permute(2, 3, 1, 0).contiguous(), table_k.permute(2, 3, 1, 0).contiguous()
    rt = rel_idx.int().permute(1, 0).contiguous()
    lib().dot_prod_with_idx_all_forward_cuda_launcher(N, M, h, d, int(n_max), L, _p(qt), _p(_i(index_q)), _p(_i(index_q_offsets)), _p(kt),
                                                      _p(_i(index_k)), _p(tq), _p(tk), _p(rt), _p(out))
    torch.cuda.synchronize()
    return out.permute(1, 0).contiguous()


def dot_prod_with_idx_all_backward(grad_out, q, k, index_q, index_q_offsets, index_k, index_k_offsets, table_q, table_k, rel_idx, n_max):
    """sptr/functional.py:294-338: dot_prod_with_idx_backward_cuda + attention_step1_backward_cuda, summed."""
    N, h, d = q.shape
    M, L = grad_out.shape[0], table_q.shape[0]
    g = grad_out.contiguous()
    gq, gk = torch.zeros_like(q), torch.zeros_like(k)
    gtq, gtk = torch.zeros_like(table_q), torch.zeros_like(table_k)
    lib().dot_prod_with_idx_backward_cuda_launcher(N, M, h, d, int(n_max), L, _p(g), _p(q), _p(_i(index_q_offsets)), _p(k), _p(_i(index_k_offsets)),
                                                   _p(_i(index_k)), _p(table_q), _p(table_k), _p(rel_idx.int().contiguous()), _p(gq), _p(gk),
                                                   _p(gtq), _p(gtk))
    gq2, gk2 = torch.zeros_like(q), torch.zeros_like(k)
    lib().attention_step1_backward_cuda_launcher(N, M, h, d, ctypes.c_uint(int(n_max)), _p(g), _p(_i(index_q)), _p(_i(index_q_offsets)), _p(_i(index_k)),
                                                 _p(_i(index_k_offsets)), _p(q), _p(k), _p(gq2), _p(gk2))
    torch.cuda.synchronize()
    return gq + gq2, gk + gk2, gtq, gtk


def attention_step2_with_rel_pos_value_forward(attn, v, index0_offsets, index1, table, rel_idx, n_max):
    """sptr/functional.py:342-372."""
    M, h = attn.shape
    N, d = index0_offsets.shape[0] - 1, v.shape[2]
    out = torch.zeros(N, h, d, dtype=torch.float32, device=v.device)
    lib().attention_step2_with_rel_pos_value_forward_cuda_launcher(N, M, h, d, int(n_max), _p(attn.contiguous()), _p(v), _p(_i(index0_offsets)),
                                                                   _p(_i(index1)), _p(table), _p(rel_idx.int().contiguous()), _p(out))
    torch.cuda.synchronize()
    return out


def attention_step2_with_rel_pos_value_backward(grad_out, attn, v, index0, index0_offsets, index1, index1_offsets, table, rel_idx, n_max):
    """sptr/functional.py:374-403."""
    N, h, d = grad_out.shape
    M, L = attn.shape[0], table.shape[0]
    ga, gv, gt = torch.zeros_like(attn), torch.zeros_like(v), torch.zeros_like(table)
    tt = table.permute(2, 3, 1, 0).contiguous()
    vt = v.permute(1, 2, 0).contiguous()
    rt = rel_idx.int().permute(1, 0).contiguous()
    lib().attention_step2_with_rel_pos_value_backward_cuda_launcher(N, M, h, d, L, int(n_max), _p(grad_out.contiguous()), _p(_i(index0)),
                                                                    _p(_i(index0_offsets)), _p(_i(index1)), _p(_i(index1_offsets)),
                                                                    _p(attn.contiguous()), _p(vt), _p(tt), _p(rt), _p(ga), _p(gv), _p(gt))
    torch.cuda.synchronize()
    return ga, gv, gt
