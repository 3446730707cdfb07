"""sptr.functional (third_party/SparseTransformer/sptr/functional.py): the step operators with the reference's
signatures, and `window_attention`, the fused operator the modules actually run.

The step operators (attention_step1 / dot_prod_with_idx / dot_prod_with_idx_all / attention_step2 /
attention_step2_with_rel_pos_value) keep the reference's call signatures for code that composes them by hand; they are
thin device-side index arithmetic (torch gathers / index_add on the GPU, autograd for the backward) because nothing on
the model path calls them any more: sparse_self_attention goes through `window_attention`, ONE CUDA kernel per direction
that never materialises the [M, h] score matrix (csrc/window_attn.cu)."""
import torch
from torch.autograd import Function

from .. import ops
from .._lib import check, lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("u2mkd_b200.sptr ops need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))


def window_offsets(counts: torch.Tensor):
    """(win_off int32 [n+1], sq_off int32 [n+1]) of window-sorted points: first row / first pair of every window."""
    c = counts.long()
    z = c.new_zeros(1)
    return torch.cat([z, c.cumsum(0)]).int(), torch.cat([z, (c * c).cumsum(0)]).int()


def precompute_all(N, n, n_max, counts):
    """sptr.precompute_all (sptr/functional.py:146-170): counts int [n] -> index_0_offsets [N+1], index_1_offsets [N],
    index_0 [M], index_1 [M] (int32), pair m = sq_off[w] + i * len + t <-> (query start + i, key start + t).  One host read
    (M sizes the outputs), as in the reference."""
    _need_cuda(counts)
    counts = counts.contiguous()
    win_off, sq_off = window_offsets(counts)
    M = int(sq_off[-1].item())
    dev = counts.device
    index_0_offsets = torch.empty(N + 1, dtype=torch.int32, device=dev)
    index_1_offsets = torch.empty(N, dtype=torch.int32, device=dev)
    index_0 = torch.empty(M, dtype=torch.int32, device=dev)
    index_1 = torch.empty(M, dtype=torch.int32, device=dev)
    check(lib().u2_window_pairs(win_off.data_ptr(), sq_off.data_ptr(), int(n), index_0_offsets.data_ptr(),
                                index_1_offsets.data_ptr(), index_0.data_ptr(), index_1.data_ptr(), ops._st()))
    ops._count()
    index_0_offsets[N] = M
    # the fused operator needs the window boundaries, not the per-pair lists: they ride along on the offsets tensor
    index_0_offsets._u2_windows = (win_off, sq_off, int(n))
    return index_0_offsets, index_1_offsets, index_0, index_1


# ------------------------------------------------------------------ step operators (reference signatures)
def _tsum(table, rel_idx):
    r = rel_idx.long()
    return table[r[:, 0], 0] + table[r[:, 1], 1] + table[r[:, 2], 2]


def attention_step1(q, k, index0, index0_offsets, index1, index1_offsets, n_max):
    """sptr.attention_step1 (sptr/functional.py:9-79): [M, h] content scores."""
    _need_cuda(q, k, index0, index1)
    return (q[index0.long()] * k[index1.long()]).sum(-1)


def dot_prod_with_idx(q, index_q, index_q_offsets, n_max, k, index_k_offsets, index_k, table_q, table_k, rel_idx):
    """sptr.dot_prod_with_idx (sptr/functional.py:172-251): relative-position scores only."""
    _need_cuda(q, k, index_q, index_k, table_q, table_k, rel_idx)
    return (q[index_q.long()] * _tsum(table_q, rel_idx)).sum(-1) + (k[index_k.long()] * _tsum(table_k, rel_idx)).sum(-1)


def dot_prod_with_idx_all(q, index_q, index_q_offsets, k, index_k, index_k_offsets, table_q, table_k, rel_idx, n_max):
    """sptr.dot_prod_with_idx_all (sptr/functional.py:253-340): content + relative-position scores."""
    return attention_step1(q, k, index_q, index_q_offsets, index_k, index_k_offsets, n_max) + \
        dot_prod_with_idx(q, index_q, index_q_offsets, n_max, k, index_k_offsets, index_k, table_q, table_k, rel_idx)


def attention_step2(attn, v, index0, index0_offsets, index1, index1_offsets, n_max):
    """sptr.attention_step2 (sptr/functional.py:81-144)."""
    _need_cuda(attn, v, index0, index1)
    N = index0_offsets.shape[0] - 1
    out = torch.zeros((N,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
    return out.index_add(0, index0.long(), attn[:, :, None] * v[index1.long()])


def attention_step2_with_rel_pos_value(attn, v, index0, index0_offsets, n_max, index1, index1_offsets, table, rel_idx):
    """sptr.attention_step2_with_rel_pos_value (sptr/functional.py:342-405)."""
    _need_cuda(attn, v, index0, index1, table, rel_idx)
    N = index0_offsets.shape[0] - 1
    out = torch.zeros((N,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
    return out.index_add(0, index0.long(), attn[:, :, None] * (v[index1.long()] + _tsum(table, rel_idx)))


# ------------------------------------------------------------------ the fused operator
class WindowAttentionFn(Function):
    @staticmethod
    def forward(ctx, q, k, v, table_q, table_k, table_v, win_off, sq_off, n_windows, rel_idx):
        _need_cuda(q, k, v, win_off, sq_off, rel_idx, table_q)
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        N, h, d = q.shape
        rel = rel_idx is not None
        L = 0
        if rel:
            rel_idx = rel_idx.contiguous().int()
            table_q, table_k, table_v = (t.contiguous().float() for t in (table_q, table_k, table_v))
            L = table_q.shape[0]
            assert table_q.shape == table_k.shape == table_v.shape == (L, 3, h, d), (table_q.shape, (L, 3, h, d))
        if not lib().u2_window_attn_supported(d, L):
            raise RuntimeError(f"window_attention: head_dim {d} / table length {L} not supported (head_dim 16 or 32, L <= 64)")
        out = torch.empty_like(q)
        lse = torch.empty((N, h), dtype=torch.float32, device=q.device)
        check(lib().u2_window_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), win_off.data_ptr(), sq_off.data_ptr(),
                                       int(n_windows), h, d, ops._ptr(rel_idx), ops._ptr(table_q if rel else None),
                                       ops._ptr(table_k if rel else None), ops._ptr(table_v if rel else None), L,
                                       out.data_ptr(), lse.data_ptr(), ops._st()))
        ops._count()
        ctx.save_for_backward(q, k, v, table_q if rel else None, table_k if rel else None, table_v if rel else None,
                              win_off, sq_off, rel_idx, out, lse)
        ctx.n_windows = int(n_windows)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, tq, tk, tv, win_off, sq_off, rel_idx, out, lse = ctx.saved_tensors
        N, h, d = q.shape
        dout = dout.contiguous().float()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        rel = rel_idx is not None
        dtq = torch.empty_like(tq) if rel else None
        dtk = torch.empty_like(tk) if rel else None
        dtv = torch.empty_like(tv) if rel else None
        L = tq.shape[0] if rel else 0
        sbytes = lib().u2_window_attn_bwd_scratch_bytes(ctx.n_windows, h, d, L)
        scratch = ops._ws("wattn", sbytes, q.device) if sbytes else None
        check(lib().u2_window_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), win_off.data_ptr(), sq_off.data_ptr(),
                                       ctx.n_windows, h, d, ops._ptr(rel_idx), ops._ptr(tq), ops._ptr(tk), ops._ptr(tv),
                                       L, out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
                                       dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ops._ptr(dtq), ops._ptr(dtk), ops._ptr(dtv),
                                       ops._ptr(scratch), scratch.numel() if scratch is not None else 0, ops._st()))
        ops._count(2 if rel else 1)
        return dq, dk, dv, dtq, dtk, dtv, None, None, None, None


def window_attention(q, k, v, win_off, sq_off, n_windows, rel_idx=None, table_q=None, table_k=None, table_v=None):
    """softmax_j(q_i.k_j + q_i.Tq[r_ij] + k_j.Tk[r_ij]) (v_j + Tv[r_ij]) over the keys j of query i's window, for
    window-sorted q / k / v [N, h, head_dim] (q already scaled): the whole of sparse_self_attention between its sort and
    un-sort (sptr/modules.py:36-62) as one kernel.  T[r_ij] = T[r0, 0] + T[r1, 1] + T[r2, 2] with (r0, r1, r2) = rel_idx[m],
    m = sq_off[w] + i * n_w + j; rel_idx / tables None = no relative position encoding."""
    return WindowAttentionFn.apply(q, k, v, table_q, table_k, table_v, win_off, sq_off, n_windows, rel_idx)
