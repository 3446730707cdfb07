"""Tensor-level bindings of the C-ABI (include/u2mkd.h): torch CUDA tensors in, raw device
pointers + the current stream down, autograd Functions on top.

This is the whole "torch extension": torch supplies device memory, streams and autograd
plumbing; all arithmetic of the path runs in libu2mkd_b200.so.  CUDA tensors only — a CPU
tensor raises (there is deliberately no CPU path in the product; the CPU restatement lives
in oracle/ and is test infrastructure).
"""
from __future__ import annotations

import ctypes
import threading
from typing import Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import MATH_BF16, MATH_FP32, MATH_TF32, check, lib

MATH_BF16X3 = 3  # host-level mode: fp32-grade results from three bf16 tensor-core products (ConvolutionX3Fn)
_MATH_NAMES = {"fp32": MATH_FP32, "tf32": MATH_TF32, "bf16": MATH_BF16, "bf16x3": MATH_BF16X3}
_state = {"math": MATH_FP32, "sort_tiles": True, "overlap_rows": 1 << 30}

# instrumentation used by bench.py: number of kernels this library launched, and an optional
# per-launch CUDA-event timer for the conv kernels (the dominant kernel of the path)
stats = {"launches": 0}
conv_timer = None  # object with .record(kind, table, n_dst, K, c_src, c_dst, ev_start, ev_stop)


_count_lock = threading.Lock()


def _count(n: int = 1) -> None:
    with _count_lock:   # the coordinate prefetch launches from a worker thread
        stats["launches"] += n


def set_math(mode: str) -> None:
    """Arithmetic of the conv GEMMs: 'fp32' (FFMA, parity mode), 'tf32' or 'bf16' (tcgen05), or 'bf16x3' (tcgen05, fp32-grade:
    operands split into bf16 hi + lo, three products accumulated in fp32 — rel ~1e-5, the tensor-core parity mode)."""
    m = _MATH_NAMES[mode]
    if m != MATH_FP32 and not lib().u2_has_tensor_core_path():
        raise RuntimeError("this build of libu2mkd_b200.so has no tcgen05 conv path")
    _state["math"] = m


def set_sort_tiles(flag: bool) -> None:
    """Mask-sorted tile order for the tensor-core conv (on by default)."""
    _state["sort_tiles"] = bool(flag)


def get_math() -> str:
    return {v: k for k, v in _MATH_NAMES.items()}[_state["math"]]


def _need_cuda(*tensors):
    """Operands must live on the CURRENT CUDA device: kernels are launched on its current stream (`_st`)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("u2mkd_b200 ops need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"u2mkd_b200 ops launch on the current device cuda:{cur}, but an operand lives on {t.device}; "
                               "wrap the call in torch.cuda.device(tensor.device)")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _st():
    """cudaStream_t of torch's current stream (raw getter: ~20x cheaper than current_stream())."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_workspaces = {}


def _ws(name: str, nbytes: int, device) -> torch.Tensor:
    """Cached scratch buffer (uint8, >= nbytes) for kernels that only need it while they run.  Reuse is ordered by the
    stream: every user of one `name` launches on the same stream, so the next kernel that overwrites the buffer runs after
    the previous reader (forward: the caller's stream; autograd replays backward on that stream too).  Side-stream users
    pass their own name.  Saves ~350 torch.empty calls (1.5 ms of host time) per training step."""
    key = (name, torch.device(device).index, _st())  # per stream: the coordinate prefetch runs the map builders on its own
    t = _workspaces.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = t
    return t


# -------------------------------------------------------------------------------- hashing
def sphash(coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """spf.sphash (core/models/utils.py:19,43,49,86,92): int32 [N,4] (+[K,3]) -> int64 [N] / [K,N]."""
    _need_cuda(coords, offsets)
    assert coords.dtype == torch.int, coords.dtype
    assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
    coords = coords.contiguous()
    n = coords.shape[0]
    if offsets is None:
        out = torch.empty(n, dtype=torch.int64, device=coords.device)
        check(lib().u2_hash(coords.data_ptr(), n, None, 0, out.data_ptr(), _st()))
        _count()
        return out
    assert offsets.dtype == torch.int, offsets.dtype
    assert offsets.ndim == 2 and offsets.shape[1] == 3, offsets.shape
    offsets = offsets.contiguous()
    K = offsets.shape[0]
    out = torch.empty((K, n), dtype=torch.int64, device=coords.device)
    check(lib().u2_hash(coords.data_ptr(), n, offsets.data_ptr(), K, out.data_ptr(), _st()))
    _count()
    return out


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    """spf.sphashquery (core/models/utils.py:21,50,93,135): position of each query in references, -1 if absent."""
    _need_cuda(queries, references)
    assert queries.dtype == torch.long and references.dtype == torch.long
    sizes = queries.size()
    q = queries.contiguous().view(-1)
    references = references.contiguous()
    n = references.shape[0]
    tbytes = lib().u2_hash_table_bytes(n)
    table = torch.empty(tbytes, dtype=torch.uint8, device=q.device)
    st = _st()
    check(lib().u2_hash_table_build(references.data_ptr(), n, table.data_ptr(), tbytes, st))
    out = torch.empty(q.shape[0], dtype=torch.int64, device=q.device)
    check(lib().u2_hash_table_query(table.data_ptr(), tbytes, q.data_ptr(), q.shape[0], out.data_ptr(), st))
    _count(2)
    return out.view(*sizes)


def spcount(coords: torch.Tensor, num: int) -> torch.Tensor:
    """spf.spcount (core/models/utils.py:22,51)."""
    _need_cuda(coords)
    assert coords.dtype == torch.int
    coords = coords.contiguous()
    out = torch.empty(int(num), dtype=torch.int, device=coords.device)
    check(lib().u2_count(coords.data_ptr(), coords.shape[0], out.data_ptr(), int(num), _st()))
    _count()
    return out


def coord_table(coords: torch.Tensor) -> torch.Tensor:
    """Hash table of a coordinate set (int32 [n,4] -> row index), cached on the tensor: one build per tensor stride
    serves the kernel maps, point_to_voxel and voxel_to_point of that stride (the coordinate tensors are shared through
    SparseTensor.cmaps, so the cache follows them)."""
    _need_cuda(coords)
    assert coords.dtype == torch.int and coords.ndim == 2 and coords.shape[1] == 4 and coords.is_contiguous()
    st = getattr(coords, "_u2_table", None)
    if st is not None and st[1] == coords._version and st[2] == coords.data_ptr():
        return st[0]
    n = coords.shape[0]
    tbytes = lib().u2_hash_table_bytes(n)
    table = torch.empty(tbytes, dtype=torch.uint8, device=coords.device)
    check(lib().u2_coord_table_build(coords.data_ptr(), n, table.data_ptr(), tbytes, _st()))
    _count(2)
    coords._u2_table = (table, coords._version, coords.data_ptr())
    return table


def coord_query(queries: torch.Tensor, ref_coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row of every query coordinate (+ each of the K offsets) in ref_coords, -1 if absent: int64 [N] / [K, N].  Same
    result as spf.sphashquery(spf.sphash(queries[, offsets]), spf.sphash(ref_coords)) (core/models/utils.py:49-50,86-93)
    without the intermediate hash tensors and with the cached table of ref_coords."""
    _need_cuda(queries, ref_coords, offsets)
    assert queries.dtype == torch.int and queries.ndim == 2 and queries.shape[1] == 4
    queries = queries.contiguous()
    table = coord_table(ref_coords.contiguous() if not ref_coords.is_contiguous() else ref_coords)
    n = queries.shape[0]
    K = 1 if offsets is None else offsets.shape[0]
    if offsets is not None:
        offsets = offsets.contiguous().int()
    out = torch.empty((K, n) if offsets is not None else (n,), dtype=torch.int64, device=queries.device)
    check(lib().u2_coord_table_query(table.data_ptr(), table.numel(), queries.data_ptr(), n, _ptr(offsets), K, out.data_ptr(),
                                     _st()))
    _count()
    return out


def unique_voxelize(coords: torch.Tensor):
    """Index part of initial_voxelize (core/models/utils.py:19-25) in one call: int32 [N,4] floored coordinates ->
    (idx_query int64 [N], counts int32 [n_vox], voxel_coords int32 [n_vox,4]); voxel order = ascending FNV hash, exactly
    what torch.unique(sphash(coords)) gives the reference.  One host sync (the voxel count sizes every later tensor)."""
    _need_cuda(coords)
    assert coords.dtype == torch.int and coords.ndim == 2 and coords.shape[1] == 4, (coords.dtype, coords.shape)
    coords = coords.contiguous()
    n = coords.shape[0]
    dev = coords.device
    idx_query = torch.empty(n, dtype=torch.int64, device=dev)
    counts = torch.empty(n, dtype=torch.int, device=dev)
    vox = torch.empty((n, 4), dtype=torch.int, device=dev)
    n_vox = torch.empty(1, dtype=torch.int64, device=dev)
    sbytes = lib().u2_unique_voxelize_scratch_bytes(n)
    scratch = _ws("uvox", sbytes, dev)
    check(lib().u2_unique_voxelize(coords.data_ptr(), n, idx_query.data_ptr(), counts.data_ptr(), vox.data_ptr(),
                                   n_vox.data_ptr(), scratch.data_ptr(), scratch.numel(), _st()))
    _count(5)
    m = int(n_vox.item())
    return idx_query, counts[:m], vox[:m]


def _unique_voxelize_launch(coords: torch.Tensor, n_dev: torch.Tensor):
    """unique_voxelize without the host read of the voxel count (written to n_dev[0]); outputs sized for n voxels."""
    coords = coords.contiguous()
    n = coords.shape[0]
    dev = coords.device
    idx_query = torch.empty(n, dtype=torch.int64, device=dev)
    counts = torch.empty(n, dtype=torch.int, device=dev)
    vox = torch.empty((n, 4), dtype=torch.int, device=dev)
    scratch = _ws("uvox", lib().u2_unique_voxelize_scratch_bytes(n), dev)
    check(lib().u2_unique_voxelize(coords.data_ptr(), n, idx_query.data_ptr(), counts.data_ptr(), vox.data_ptr(),
                                   n_dev.data_ptr(), scratch.data_ptr(), scratch.numel(), _st()))
    _count(5)
    return idx_query, counts, vox


def _downsample_launch(coords: torch.Tensor, sample_stride, n_dev: torch.Tensor) -> torch.Tensor:
    """downsample_coords without the host read of the row count (written to n_dev[0]); output sized for n rows."""
    coords = coords.contiguous()
    n = coords.shape[0]
    sbytes = lib().u2_downsample_scratch_bytes(n)
    scratch = _ws("downs", sbytes, coords.device)
    out = torch.empty((n, 4), dtype=torch.int, device=coords.device)
    check(lib().u2_downsample_coords(coords.data_ptr(), n, int(sample_stride[0]), int(sample_stride[1]),
                                     int(sample_stride[2]), out.data_ptr(), n_dev.data_ptr(), scratch.data_ptr(), sbytes,
                                     _st()))
    _count(4)
    return out


def planned_strides():
    """Tensor strides of the coordinate sets the prebuild plan's strided convs produce, ascending."""
    out = set()
    for (in_stride, ks, stride, _dil) in _plan:
        # (the stride in {1, kernel_size} case of spdownsample, the only one U2MKD's models use; others stay lazy)
        if any(v > 1 for v in stride) and all(stride[a] in (1, ks[a]) for a in range(3)):
            out.add(tuple(in_stride[a] * stride[a] for a in range(3)))
    return sorted(out)


def coords_begin(floored: torch.Tensor):
    """Phase A of a scan batch's coordinate pipeline, no host synchronisation: voxel keys -> unique (stride-1 voxels in the
    reference's order) and, straight from the same point rows, the coordinate set of every coarser stride the prebuild
    plan knows (unique(floor(c / s) * s) of the points = of the voxels = the chained spdownsample results); all row counts
    go to one pinned host buffer with one asynchronous copy.  `floored`: int32 [N,4] point coordinates."""
    _need_cuda(floored)
    strides = planned_strides() if _state.get("prebuild", True) else []
    n_dev = torch.empty(1 + len(strides), dtype=torch.int64, device=floored.device)
    idx_query, counts, vox = _unique_voxelize_launch(floored, n_dev[0:1])
    coarse = [_downsample_launch(floored, s, n_dev[i + 1:i + 2]) for i, s in enumerate(strides)]
    n_host = torch.empty(n_dev.shape[0], dtype=torch.int64, pin_memory=True)
    n_host.copy_(n_dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return {"idx_query": idx_query, "counts": counts, "vox": vox, "strides": strides, "coarse": coarse, "n_host": n_host,
            "n_dev": n_dev, "ready": ev}


def coords_finish(h, SparseTensor):
    """Phase B: read the row counts (ONE wait, for work queued in phase A), slice the coordinate sets and build every
    planned kernel map / tile sort / pair list on them.  Returns (idx_query, counts, coords, cmaps, kmaps)."""
    h["ready"].synchronize()
    ns = [int(v) for v in h["n_host"].tolist()]
    if min(ns) < 0:
        raise RuntimeError("downsample_coords: coordinates outside [0, 2^18) or batch outside [0, 1024)")
    coords = h["vox"][:ns[0]]
    holder = SparseTensor(coords.new_zeros((0, 1), dtype=torch.float32), coords, 1)
    holder.cmaps[holder.stride] = coords
    for s, c, m in zip(h["strides"], h["coarse"], ns[1:]):
        holder.cmaps[s] = c[:m]
    prebuild_maps(holder)
    return h["idx_query"], h["counts"][:ns[0]], coords, holder.cmaps, holder.kmaps


# -------------------------------------------------------------------------------- voxelize
class VoxelizeFn(Function):
    """spf.spvoxelize (core/models/utils.py:24,26,58): scatter-mean with autograd."""

    @staticmethod
    def forward(ctx, feats, coords, counts):
        _need_cuda(feats, coords, counts)
        in_dtype = feats.dtype
        feats = feats.contiguous().float()
        coords = coords.contiguous().int()
        counts = counts.contiguous().int()
        n_pts, c = feats.shape
        n_vox = counts.shape[0]
        out = torch.empty((n_vox, c), dtype=torch.float32, device=feats.device)
        check(lib().u2_voxelize_fwd(feats.data_ptr(), n_pts, c, coords.data_ptr(), counts.data_ptr(), out.data_ptr(),
                                    n_vox, _st()))
        _count()
        ctx.for_backwards = (coords, counts, n_pts, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        coords, counts, n_pts, in_dtype = ctx.for_backwards
        g = grad_output.contiguous().float()
        n_vox, c = g.shape
        gin = torch.empty((n_pts, c), dtype=torch.float32, device=g.device)
        check(lib().u2_voxelize_bwd(g.data_ptr(), n_vox, c, coords.data_ptr(), counts.data_ptr(), gin.data_ptr(), n_pts,
                                    _st()))
        _count()
        return gin.to(in_dtype), None, None


def spvoxelize(feats, coords, counts):
    return VoxelizeFn.apply(feats, coords, counts)


# -------------------------------------------------------------------------------- devoxelize
def calc_ti_weights(coords: torch.Tensor, idx_query: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """spf.calc_ti_weights (core/models/utils.py:94): fp32 [N,4], int64 [8,N] -> fp32 [8,N]."""
    _need_cuda(coords, idx_query)
    with torch.no_grad():
        coords = coords.contiguous().float()
        assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
        idx_query = idx_query.contiguous().long()
        n = coords.shape[0]
        assert idx_query.shape == (8, n), idx_query.shape
        w = torch.empty((8, n), dtype=torch.float32, device=coords.device)
        check(lib().u2_ti_weights(coords.data_ptr(), idx_query.data_ptr(), n, float(scale), w.data_ptr(), _st()))
        _count()
    return w


class DevoxelizeFn(Function):
    """spf.spdevoxelize (core/models/utils.py:99,111)."""

    @staticmethod
    def forward(ctx, feats, coords, weights):
        _need_cuda(feats, coords, weights)
        in_dtype = feats.dtype
        feats = feats.contiguous().float()
        coords = coords.contiguous().int()
        weights = weights.contiguous().float()
        n_vox, c = feats.shape
        n_pts = coords.shape[0]
        assert coords.shape == (n_pts, 8) and weights.shape == (n_pts, 8)
        out = torch.empty((n_pts, c), dtype=torch.float32, device=feats.device)
        check(lib().u2_devoxelize_fwd(feats.data_ptr(), n_vox, c, coords.data_ptr(), weights.data_ptr(), n_pts,
                                      out.data_ptr(), _st()))
        _count()
        ctx.for_backwards = (coords, weights, n_vox, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        coords, weights, n_vox, in_dtype = ctx.for_backwards
        g = grad_output.contiguous().float()
        n_pts, c = g.shape
        gf = torch.empty((n_vox, c), dtype=torch.float32, device=g.device)
        check(lib().u2_devoxelize_bwd(g.data_ptr(), n_pts, c, coords.data_ptr(), weights.data_ptr(), gf.data_ptr(), n_vox,
                                      _st()))
        _count()
        return gf.to(in_dtype), None, None


def spdevoxelize(feats, coords, weights):
    return DevoxelizeFn.apply(feats, coords, weights)


# -------------------------------------------------------------------------------- kernel maps
def downsample_coords(coords: torch.Tensor, sample_stride: Tuple[int, int, int]) -> torch.Tensor:
    """Unique rows of (floor(c / s) * s, b), sorted by (b, x, y, z); one host sync for the row count."""
    _need_cuda(coords)
    assert coords.dtype == torch.int and coords.ndim == 2 and coords.shape[1] == 4
    coords = coords.contiguous()
    n = coords.shape[0]
    sbytes = lib().u2_downsample_scratch_bytes(n)
    scratch = torch.empty(sbytes, dtype=torch.uint8, device=coords.device)
    out = torch.empty((n, 4), dtype=torch.int, device=coords.device)
    n_out = torch.empty(1, dtype=torch.int64, device=coords.device)
    check(lib().u2_downsample_coords(coords.data_ptr(), n, int(sample_stride[0]), int(sample_stride[1]),
                                     int(sample_stride[2]), out.data_ptr(), n_out.data_ptr(), scratch.data_ptr(), sbytes,
                                     _st()))
    _count(4)
    m = int(n_out.item())
    if m < 0:
        raise RuntimeError("downsample_coords: coordinates outside [0, 2^18) or batch outside [0, 1024)")
    return out[:m]


# ------------------------------------------------------------------ kernel-map prebuild plan
# Kernel maps depend on coordinates only.  Built lazily (as torchsparse does, inside the first conv that needs one) their
# host synchronisations — the row count of every strided coordinate set — sit in the middle of the forward pass, where the
# host then stops running ahead of the GPU (measured: forward 15.1 ms wall for 10.6 ms of host enqueue time and ~11.5 ms
# of kernels).  The first forward pass of a process therefore records which maps / derived tables it needed, in order;
# later passes rebuild exactly those right after initial_voxelize (models.initial_voxelize -> prebuild_maps), so every
# synchronisation happens before the first feature kernel is queued and the rest of the step is enqueued without a stall.
_plan = {}       # plan_key -> set of derived tables used ("sortF", "sortT", "flat"), insertion-ordered


def _plan_note(plan_key, what=None) -> None:
    if plan_key is None or not _state.get("prebuild", True):
        return
    flags = _plan.setdefault(plan_key, set())
    if what is not None:
        flags.add(what)


def set_prebuild(flag: bool) -> None:
    """Rebuild the recorded kernel maps at the start of every forward pass (default on); off = lazily, as the reference."""
    _state["prebuild"] = bool(flag)
    if not flag:
        _plan.clear()


def prebuild_maps(x) -> int:
    """Build, for the coordinate set of the stride-1 SparseTensor `x`, every kernel map (and mask-sorted table / pair list)
    that an earlier forward pass asked for; results land in x.cmaps / x.kmaps, which every later tensor shares by reference
    (core/models/utils.py:60-61), so the convs find them as cache hits.  Returns the number of maps built."""
    if not _plan or not _state.get("prebuild", True):
        return 0
    from .torchsparse.nn.functional import build_map_for  # late import: functional imports ops
    built = 0
    for key, flags in list(_plan.items()):
        if key in x.kmaps or x.cmaps.get(key[0]) is None:
            continue
        kmap = build_map_for(x.cmaps, x.kmaps, key)
        built += 1
        if "sortF" in flags:
            kmap.sorted_tables(False)
        if "sortT" in flags:
            kmap.sorted_tables(True)
        if "flat" in flags:
            kmap.flat_pairs
            kmap.dense_hint()
    return built


class CoordPrefetch:
    """Coordinate-only work of the NEXT scan batch (voxel keys, unique, coarse coordinate sets, kernel maps, tile sorts, pair
    lists — everything that depends on coordinates alone) on a high-priority side stream, in two phases queued from the
    training loop's own thread:
        begin(fn)   before the current step is queued: fn() launches the counting kernels (no host wait); they run as soon
                    as the previous step has finished on the GPU, next to the start of the current one;
        finish(fn)  after the current step is queued: fn() reads the counts — long since there, the host is several ms
                    ahead of the GPU by then — and queues the map builders, which run next to the step's tail.
    The consumer's stream waits for the event finish() returns.  Neither phase blocks in steady state, so no worker thread
    is needed (profiles/r2_prefetch_steplog.md).

    Memory: tensors allocated on the side stream are consumed on the main stream, which the caching allocator does not
    track.  Results are therefore kept alive here for one more generation, and begin() orders the side stream behind the
    main-stream position at that moment before it releases the older generation: the step that used it was queued before,
    so a block can only be reused after its last reader."""

    def __init__(self):
        self.stream = None
        self.keep = []

    def _side(self):
        main = torch.cuda.current_stream()
        if self.stream is None or self.stream.device != main.device:
            self.stream = torch.cuda.Stream(device=main.device, priority=-1)
            self.keep = []
        return main, self.stream

    def reset(self):
        """Forget the kept generations.  Only after a device synchronisation (nothing queued can still read them)."""
        self.keep = []

    def begin(self, fn):
        main, side = self._side()
        mark = torch.cuda.Event()
        mark.record(main)
        side.wait_event(mark)        # inputs exist, and the steps that used older results are over ...
        del self.keep[:-1]           # ... so those may go back to the side stream's pool (the newest is about to be used)
        with torch.cuda.stream(side):
            return fn()

    def finish(self, fn):
        _main, side = self._side()
        with torch.cuda.stream(side):
            out = fn()
            done = torch.cuda.Event()
            done.record(side)
        self.keep.append(out)
        return out, done


coord_prefetch = CoordPrefetch()


class KernelMap:
    """Kernel map of one sparse Conv3d: dense neighbour tables on the device.

    nbr  int32 [K, ld_out]: nbr[k, o]  = input row matching output row o at offset k, or -1
    nbrT int32 [K, ld_in ]: nbrT[k, i] = output row of that pair, or -1
    Indexing as a 3-list reproduces the reference contract read at
    core/models/sphereformer/unet_spherical_transformer.py:226-232:
    [nbmaps int64 [M,2] rows (in, out) ordered by (k, out), nbsizes [K], (N_in, N_out)].
    """

    def __init__(self, nbr, nbrT, nbsizes, n_in, n_out, offsets_host=None, same_coords=False):
        self.nbr, self.nbrT, self.nbsizes = nbr, nbrT, nbsizes
        self.n_in, self.n_out = n_in, n_out
        self.K = nbr.shape[0]
        self.offsets_host = offsets_host  # [K][3] python ints (host), for the sort key layout
        self.same_coords = same_coords    # submanifold map: input and output rows are the same voxels
        self._nbmaps = None
        self._flat = None
        self._sorted = {}
        self.plan_key = None  # (tensor_stride, kernel_size, stride, dilation) when built by F.conv3d: feeds the prebuild plan

    def _bitpos(self):
        """Bit of each offset in the row sort key: rare offsets (corners, then edges, vertical
        before horizontal) in the high bits, the always-present centre lowest."""
        K = self.K
        offs = self.offsets_host if self.offsets_host is not None else [[0, 0, 0]] * K
        order = sorted(range(K), key=lambda k: (sum(1 for v in offs[k] if v != 0), offs[k][2] != 0, k))
        pos = [0] * K
        for rank, k in enumerate(order):
            pos[k] = rank
        return (ctypes.c_int32 * K)(*pos)

    def sorted_tables(self, transposed_side: bool):
        """(tableP [K, ld], perm [ld]) for the tensor-core conv: rows sorted by neighbour mask.
        transposed_side=False: the nbr table (rows = output voxels); True: nbrT (rows = input voxels)."""
        key = bool(transposed_side)
        if key not in self._sorted:
            _plan_note(self.plan_key, "sortT" if key else "sortF")
            table = self.nbrT if key else self.nbr
            n_rows = self.n_in if key else self.n_out
            ld = table.shape[1]
            tabP = torch.empty_like(table)
            tmask = torch.empty(ld // _PAD, dtype=torch.int, device=table.device)
            other = self._sorted.get(not key)
            st = _st()
            if self.same_coords and other is not None:
                # nbrT[k][i] == nbr[K-1-k][i] on a submanifold map: the same row order is as good
                perm = other[1]
                check(lib().u2_kmap_sort_rows(table.data_ptr(), ld, n_rows, self.K, None, perm.data_ptr(), None,
                                              tabP.data_ptr(), tmask.data_ptr(), None, 0, st))
                _count(2)
            else:
                perm = torch.empty(ld, dtype=torch.int, device=table.device)
                sbytes = lib().u2_kmap_sort_scratch_bytes(n_rows)
                scratch = torch.empty(sbytes, dtype=torch.uint8, device=table.device)
                check(lib().u2_kmap_sort_rows(table.data_ptr(), ld, n_rows, self.K, self._bitpos(), None, perm.data_ptr(),
                                              tabP.data_ptr(), tmask.data_ptr(), scratch.data_ptr(), sbytes, st))
                _count(8)
            self._sorted[key] = (tabP, perm, tmask)
        return self._sorted[key]

    def dense_hint(self):
        """(dense_k, flag): offset of a submanifold / identity map that pairs every row with itself (the centre tap; the only
        offset of a 1x1x1 layer) and a device int32 flag saying that it really does (a coordinate set with duplicates would
        not) — the wgrad kernel streams that offset's rows with TMA tile loads.  (-1, None) for strided / transposed maps."""
        if getattr(self, "_dense", None) is None:
            k = -1
            if self.same_coords and self.n_in == self.n_out and self.offsets_host is not None:
                for i, o in enumerate(self.offsets_host):
                    if all(v == 0 for v in o):
                        k = i
            flag = None
            if k >= 0:
                n = self.n_out
                flag = (self.nbr[k, :n] == torch.arange(n, dtype=torch.int, device=self.nbr.device)).all().to(torch.int32).view(1)
                _count(3)
            self._dense = (k, flag)
        return self._dense

    @property
    def flat_pairs(self) -> torch.Tensor:
        """int32 flat indices k*ld_out + out_row of the valid entries of nbr, ascending (device-side
        compaction, no host sync; sized for the worst case, valid prefix = nbsizes.sum())."""
        if self._flat is None:
            _plan_note(self.plan_key, "flat")
            total = self.nbr.numel()
            # a submanifold / strided map has at most min(K * n_out, K * n_in) pairs
            cap = self.K * min(self.n_out, self.n_in) if self.K * min(self.n_out, self.n_in) > 0 else 1
            flat = torch.empty(cap, dtype=torch.int, device=self.nbr.device)
            sbytes = lib().u2_kmap_pairs_scratch_bytes(total)
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=self.nbr.device)
            check(lib().u2_kmap_pairs(self.nbr.data_ptr(), total, flat.data_ptr(), scratch.data_ptr(), sbytes, _st()))
            _count()
            self._flat = flat
        return self._flat

    @property
    def nbmaps(self) -> torch.Tensor:
        if self._nbmaps is None:
            res = self.nbr[:, :self.n_out]
            nz = torch.nonzero(res != -1)
            nz[:, 0] = res[nz[:, 0], nz[:, 1]].long()
            self._nbmaps = nz
        return self._nbmaps

    @property
    def sizes(self):
        return (self.n_in, self.n_out)

    def __getitem__(self, i):
        return (self.nbmaps, self.nbsizes, self.sizes)[i]

    def __len__(self):
        return 3

    def __iter__(self):
        return iter((self.nbmaps, self.nbsizes, self.sizes))


_PAD = 128


def _pad(n: int) -> int:
    return max(_PAD, (n + _PAD - 1) // _PAD * _PAD)


def build_kernel_map(in_coords: torch.Tensor, out_coords: torch.Tensor, offsets: torch.Tensor,
                     offsets_host=None) -> KernelMap:
    """For every output voxel o and offset k: the input voxel at coord(o) + offset_k."""
    _need_cuda(in_coords, out_coords, offsets)
    same = in_coords is out_coords
    if offsets_host is None:
        offsets_host = offsets.cpu().tolist()
    in_coords = in_coords.contiguous()
    out_coords = out_coords.contiguous()
    offsets = offsets.contiguous().int()
    dev = in_coords.device
    n_in, n_out, K = in_coords.shape[0], out_coords.shape[0], offsets.shape[0]
    ld_in, ld_out = _pad(n_in), _pad(n_out)
    nbr = torch.empty((K, ld_out), dtype=torch.int, device=dev)
    nbrT = torch.empty((K, ld_in), dtype=torch.int, device=dev)
    nbsizes = torch.empty(K, dtype=torch.int, device=dev)
    table = coord_table(in_coords)  # built once per coordinate set (scratch_bytes = 0: pre-built table)
    check(lib().u2_kmap_build(in_coords.data_ptr(), n_in, out_coords.data_ptr(), n_out, offsets.data_ptr(), K,
                              nbr.data_ptr(), ld_out, nbrT.data_ptr(), ld_in, nbsizes.data_ptr(), table.data_ptr(),
                              0, _st()))
    _count(1)
    return KernelMap(nbr, nbrT, nbsizes, n_in, n_out, offsets_host, same)


_identity_maps = {}


def identity_kernel_map(n: int, device) -> KernelMap:
    """Kernel map of a 1x1x1 conv / a Linear layer over n rows: one offset, row i <-> row i.  With it the dense layers
    (ResidualBlock shortcut convs core/models/build_blocks.py:74-78, point MLPs core/models/semantickitti/spvcnn.py:58-76)
    run on the same tcgen05 kernels — and the same fused BatchNorm epilogue — as the k=3 convs instead of cuBLAS.
    Depends on n only: a small cache serves every layer of a step."""
    key = (int(n), torch.device(device).index)
    km = _identity_maps.get(key)
    if km is None:
        if len(_identity_maps) >= 32:
            _identity_maps.pop(next(iter(_identity_maps)))
        ld = _pad(n)
        nbr = torch.arange(ld, dtype=torch.int, device=device)
        nbr[n:] = -1
        nbr = nbr.view(1, ld)
        nbsizes = torch.full((1,), n, dtype=torch.int, device=device)
        km = KernelMap(nbr, nbr, nbsizes, n, n, [[0, 0, 0]], True)
        km._sorted = {False: (nbr, nbr[0], None), True: (nbr, nbr[0], None)}  # every row has the same mask: nothing to sort
        km._flat = nbr[0]                                                      # flat pair list k * ld + row = row
        km._dense = (0, torch.ones(1, dtype=torch.int32, device=device))       # every pair is (row j, row j)
        _count(2)
        _identity_maps[key] = km
    return km


def dense_tc_supported(cin: int, cout: int) -> bool:
    """bf16 tcgen05 kernels cover a dense [n, cin] x [cin, cout] layer (fwd, dgrad, wgrad) in the current math mode."""
    l = lib()
    return bool(_state["math"] == MATH_BF16 and _state.get("dense_tc", True)
                and l.u2_conv_tc_shape_supported(cin, cout, 1, MATH_BF16) and l.u2_conv_tc_shape_supported(cout, cin, 1, MATH_BF16)
                and l.u2_conv_wgrad_pairs_supported(cin, cout, 1, MATH_BF16))


class _ZeroGradFor(Function):
    """y = x, and a zero gradient for `param`: the bias of a Linear in front of a training-mode BatchNorm cancels out of
    the output, its exact gradient is 0 (the BatchNorm backward sums to zero over the rows)."""

    @staticmethod
    def forward(ctx, x, param):
        ctx.shape, ctx.dtype, ctx.device = param.shape, param.dtype, param.device
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g, torch.zeros(ctx.shape, dtype=ctx.dtype, device=ctx.device)


def linear_bn_relu(x, linear: torch.nn.Linear, bn, relu: bool):
    """Sequential(Linear, BatchNorm1d[, ReLU]) of the point branch (spvcnn.py:58-76) as one fused node on the tcgen05
    kernels; None if the layer / mode is not covered (caller falls back to the torch modules)."""
    cout, cin = linear.weight.shape
    if not (bn.training and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] > 1
            and bn.momentum is not None and dense_tc_supported(cin, cout) and cout % 32 == 0):
        return None
    km = identity_kernel_map(x.shape[0], x.device)
    w = linear.weight.t().contiguous().unsqueeze(0)  # [1, cin, cout], the conv kernels' layout (autograd transposes back)
    out = sparse_conv_bn_relu(x, w, km, False, bn, relu)
    if linear.bias is not None:
        if bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():  # the statistics were taken without the bias: mean(xW + b) = mean(xW) + b
                bn.running_mean.add_(linear.bias.detach(), alpha=float(bn.momentum))
        stash = getattr(out, "_u2_bf16", None)
        out = _ZeroGradFor.apply(out, linear.bias)
        if stash is not None:
            out._u2_bf16 = (stash[0], out._version)
    return out


# -------------------------------------------------------------------------------- convolution
def _timed(kind, kmap, n_dst, K, c_src, c_dst, launch):
    """Run `launch()`; if bench.py installed a conv timer, bracket it with CUDA events on the
    launching stream."""
    _count()
    if conv_timer is None:
        return launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    conv_timer.record(kind, kmap, n_dst, K, c_src, c_dst, e0, e1)


def _conv_gather_gemm(kind, kmap, x, w, w_transposed, table, n_dst, c_dst, math, side=None, blob=None, yadd=None, algo_c_src=None):
    """One fused gather-GEMM launch.  `blob`: weights already re-tiled by u2_conv_pretile for this direction (then `w` is
    not read); `yadd`: fp32 [n_dst, c_dst] added in the epilogue (sorted-tile path only)."""
    K, ld = table.shape
    n_src, c_src = x.shape
    c_rep = c_src if algo_c_src is None else algo_c_src  # channels the conv timer credits (bf16x3 rows are 3 x as wide)
    y = torch.empty((n_dst, c_dst), dtype=torch.float32, device=x.device)
    assert (x.dtype == torch.bfloat16) == (math == MATH_BF16), (x.dtype, math)
    if blob is not None:
        scratch, sbytes, wptr = blob, blob.numel(), None
    else:
        sbytes = lib().u2_conv_scratch_bytes(n_dst, K, c_src, c_dst, math)
        scratch = _ws("pretile", sbytes, x.device) if sbytes else None
        wptr = w.data_ptr()
    if side is not None and _state["sort_tiles"] and lib().u2_conv_tc_shape_supported(c_src, c_dst, K, math):
        tabP, perm, tmask = kmap.sorted_tables(side)
        _timed(kind, kmap, n_dst, K, c_rep, c_dst, lambda: check(lib().u2_conv_fwd_perm(
            x.data_ptr(), n_src, c_src, wptr, int(w_transposed), tabP.data_ptr(), perm.data_ptr(), _ptr(yadd), ld, n_dst, K,
            c_dst, y.data_ptr(), math, _ptr(scratch), sbytes, _st())))
        return y
    _timed(kind, kmap, n_dst, K, c_rep, c_dst, lambda: check(lib().u2_conv_fwd(
        x.data_ptr(), n_src, c_src, wptr, int(w_transposed), table.data_ptr(), ld, n_dst, K, c_dst, y.data_ptr(),
        math, _ptr(scratch), sbytes, _st())))
    return y if yadd is None else y.add_(yadd)


class _WeightBlobs:
    """bf16 weight blobs of every conv parameter seen so far, re-tiled together: the first conv that finds its parameter
    changed (`_version` moved: an optimizer step, a state_dict load) re-tiles ALL registered parameters, both directions, in
    one launch (u2_conv_pretile_run) instead of two small launches per layer and step (102 per SPVCNN step).  Blobs
    persist; the job table is rebuilt only when the set of parameters (or one's storage) changes."""

    def __init__(self):
        self.entries = {}       # id(param) -> [weakref, data_ptr, shape, blob tensor, version, sbytes_fwd]
        self.plan = None        # (device table, n_jobs, n_blocks)
        self.enabled = True

    def get(self, weight: torch.Tensor):
        """(blob_fwd, blob_dgrad) for a [K, Cin, Cout] parameter; None when the arena is off."""
        if not self.enabled:
            return None
        import weakref
        e = self.entries.get(id(weight))
        if e is None or e[0]() is not weight or e[1] != weight.data_ptr() or e[2] != tuple(weight.shape):
            K, cin, cout = weight.shape
            nb = K * cin * cout * 2
            if len(self.entries) >= 4096:   # parameters that come and go (or fresh wrappers every call): not what this is for
                self.entries = {k: v for k, v in self.entries.items() if v[0]() is not None}
                if len(self.entries) >= 4096:
                    self.enabled = False
                    return None
            e = [weakref.ref(weight), weight.data_ptr(), tuple(weight.shape),
                 torch.empty(2 * nb, dtype=torch.uint8, device=weight.device), weight._version, nb]
            self.entries[id(weight)] = e
            self.plan = None                # joins the one-launch plan at the next change of the parameters
            check(lib().u2_conv_pretile(weight.data_ptr(), K, cin, cout, MATH_BF16, e[3].data_ptr(), e[3].data_ptr() + nb, _st()))
            _count(2)
        elif e[4] != weight._version:
            self._retile(weight.device)
        return e[3][:e[5]], e[3][e[5]:]

    def _retile(self, device):
        import ctypes
        import numpy as np
        l = lib()
        if self.plan is None or self.plan[3] != device:
            live = {k: e for k, e in self.entries.items() if e[0]() is not None and e[3].device == device}
            self.entries = {k: e for k, e in self.entries.items() if e[0]() is not None}
            n = 2 * len(live)
            W = np.zeros(n, np.uint64); B = np.zeros(n, np.uint64)
            K = np.zeros(n, np.int32); Cs = np.zeros(n, np.int32); Cd = np.zeros(n, np.int32); T = np.zeros(n, np.int32)
            for i, e in enumerate(live.values()):
                k, cin, cout = e[2]
                W[2 * i] = W[2 * i + 1] = e[1]
                B[2 * i], B[2 * i + 1] = e[3].data_ptr(), e[3].data_ptr() + e[5]
                K[2 * i] = K[2 * i + 1] = k
                Cs[2 * i], Cd[2 * i], T[2 * i] = cin, cout, 0
                Cs[2 * i + 1], Cd[2 * i + 1], T[2 * i + 1] = cout, cin, 1
            host = np.zeros(max(l.u2_conv_pretile_plan_bytes(n), 1), np.uint8)
            nblk = ctypes.c_int64(0)
            check(l.u2_conv_pretile_plan(n, W.ctypes.data, B.ctypes.data, K.ctypes.data, Cs.ctypes.data, Cd.ctypes.data,
                                         T.ctypes.data, MATH_BF16, host.ctypes.data, host.nbytes, ctypes.addressof(nblk)))
            self.plan = (torch.from_numpy(host).to(device), n, int(nblk.value), device)
        tab, n, nblk, _ = self.plan
        check(l.u2_conv_pretile_run(tab.data_ptr(), n, nblk, _st()))
        _count()
        for e in self.entries.values():
            w = e[0]()
            if w is not None and e[3].device == device:
                e[4] = w._version


weight_blobs = _WeightBlobs()


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy of a feature matrix (operand conversion of the bf16 conv mode)."""
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().u2_cast_bf16(x.data_ptr(), x.numel(), y.data_ptr(), _st()))
    _count()
    return y


def _layer_math(math: int, c_src: int, c_dst: int, K: int) -> int:
    """bf16 kernels exist for channel counts that fill their tiles; other layers of a bf16 run use
    the tf32 entry points (which themselves fall back to FFMA for e.g. the Cin=4 stem conv)."""
    if math == MATH_BF16 and not lib().u2_conv_tc_shape_supported(c_src, c_dst, K, MATH_BF16):
        return MATH_TF32
    return math


class ConvolutionFn(Function):
    """ConvolutionFunction of torchsparse F.conv3d, one fused kernel per
    direction instead of K gather/mm/scatter rounds."""

    @staticmethod
    def forward(ctx, feats, weight, kmap: KernelMap, transposed: bool, math: int):
        _need_cuda(feats, weight)
        in_dtype = feats.dtype
        feats = feats.contiguous().float()
        weight = weight.contiguous().float()
        K, cin, cout = weight.shape
        assert K == kmap.K and feats.shape[1] == cin
        if not transposed:
            table, n_dst = kmap.nbr, kmap.n_out
            assert feats.shape[0] == kmap.n_in, (feats.shape, kmap.sizes)
        else:
            table, n_dst = kmap.nbrT, kmap.n_in
            assert feats.shape[0] == kmap.n_out, (feats.shape, kmap.sizes)
        m_fwd = _layer_math(math, cin, cout, K)
        x_op = bf16_view(feats) if m_fwd == MATH_BF16 else feats
        arena = weight_blobs.get(weight) if (m_fwd == MATH_BF16 and isinstance(weight, torch.nn.Parameter) and
                                             _layer_math(math, cout, cin, K) == MATH_BF16) else None
        out = _conv_gather_gemm("fwd", kmap, x_op, weight, False, table, n_dst, cout, m_fwd, side=bool(transposed),
                                blob=arena[0] if arena else None)
        # the bf16 copy (half the bytes) is what wgrad needs later; otherwise keep the fp32 rows
        ctx.save_for_backward(x_op, weight, arena[1] if arena else None)
        ctx.misc = (kmap, transposed, math, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        x_op, weight, blob_dgrad = ctx.saved_tensors
        kmap, transposed, math, in_dtype = ctx.misc
        g = grad_output.contiguous().float()
        K, cin, cout = weight.shape
        grad_feats = grad_weight = None
        fwd_table = kmap.nbrT if transposed else kmap.nbr
        x_is_bf16 = x_op.dtype == torch.bfloat16
        m_dgrad = _layer_math(math, cout, cin, K) if ctx.needs_input_grad[0] else MATH_FP32
        m_wgrad = MATH_BF16 if (x_is_bf16 and lib().u2_conv_wgrad_pairs_supported(cin, cout, K, MATH_BF16)) else \
            (MATH_TF32 if math != MATH_FP32 else MATH_FP32)
        g_bf16 = cast_bf16(g) if MATH_BF16 in (m_dgrad, m_wgrad) else None
        if ctx.needs_input_grad[0]:
            bwd_table = kmap.nbr if transposed else kmap.nbrT
            grad_feats = _conv_gather_gemm("dgrad", kmap, g_bf16 if m_dgrad == MATH_BF16 else g, weight, True, bwd_table,
                                           x_op.shape[0], cin, m_dgrad, side=not transposed,
                                           blob=blob_dgrad if m_dgrad == MATH_BF16 else None).to(in_dtype)
        if ctx.needs_input_grad[1]:
            if x_is_bf16 and m_wgrad != MATH_BF16:
                x_op = x_op.float()  # (not reached for SPVCNN shapes: bf16 fwd implies bf16 wgrad)
            grad_weight = torch.empty_like(weight)
            if m_wgrad != MATH_FP32 and lib().u2_conv_wgrad_pairs_supported(cin, cout, K, m_wgrad):
                flat = kmap.flat_pairs
                gw = g_bf16 if m_wgrad == MATH_BF16 else g
                dk, dflag = kmap.dense_hint() if m_wgrad == MATH_BF16 else (-1, None)
                _timed("wgrad", kmap, g.shape[0], K, cin, cout, lambda: check(lib().u2_conv_wgrad_pairs_dense(
                    x_op.data_ptr(), cin, gw.data_ptr(), cout, kmap.nbr.data_ptr(), kmap.nbr.shape[1], kmap.n_out, K,
                    flat.data_ptr(), kmap.nbsizes.data_ptr(), int(transposed), grad_weight.data_ptr(), m_wgrad, dk, _ptr(dflag),
                    _st())))
            else:
                n_dst = g.shape[0]
                m = MATH_FP32 if m_wgrad == MATH_BF16 else m_wgrad
                sbytes = lib().u2_conv_scratch_bytes(n_dst, K, cin, cout, m)
                scratch = torch.empty(sbytes, dtype=torch.uint8, device=g.device) if sbytes else None
                _timed("wgrad", kmap, n_dst, K, cin, cout, lambda: check(lib().u2_conv_wgrad(
                    x_op.data_ptr(), x_op.shape[0], cin, g.data_ptr(), n_dst, cout, fwd_table.data_ptr(),
                    fwd_table.shape[1], K, grad_weight.data_ptr(), m, _ptr(scratch), sbytes, _st())))
        return grad_feats, grad_weight, None, None, None


def split_bf16x3(x: torch.Tensor, want3: bool = True, want_parts: bool = True):
    """fp32 [n, C] -> ([hi | lo | hi] bf16 [n, 3C], hi bf16 [n, C], lo bf16 [n, C]) with x = hi + lo to ~2^-17."""
    x = x.contiguous()
    n, c = x.shape
    out3 = torch.empty((n, 3 * c), dtype=torch.bfloat16, device=x.device) if want3 else None
    hi = torch.empty((n, c), dtype=torch.bfloat16, device=x.device) if want_parts else None
    lo = torch.empty((n, c), dtype=torch.bfloat16, device=x.device) if want_parts else None
    check(lib().u2_split_bf16x3(x.data_ptr(), n, c, _ptr(out3), _ptr(hi), _ptr(lo), _st()))
    _count()
    return out3, hi, lo


def _split_weight(w: torch.Tensor):
    """(Whi, Wlo) as fp32 tensors holding bf16-representable values: the bf16 rounding inside the weight pre-tiling is then exact."""
    whi = w.bfloat16().float()
    wlo = (w - whi).bfloat16().float()
    return whi, wlo


def bf16x3_supported(cin: int, cout: int, K: int) -> bool:
    l = lib()
    return bool(cin % 8 == 0 and cout % 8 == 0
                and l.u2_conv_tc_shape_supported(3 * cin, cout, K, MATH_BF16) and l.u2_conv_tc_shape_supported(3 * cout, cin, K, MATH_BF16)
                and l.u2_conv_wgrad_pairs_supported(cin, cout, K, MATH_BF16))


class ConvolutionX3Fn(Function):
    """Sparse conv in the 'bf16x3' mode: every fp32 operand is split as hi + lo (two bf16 numbers, 16 significant bits) and
    x.w ~ hi.Whi + lo.Whi + hi.Wlo (the lo.Wlo term is below 2^-16 relative) is computed by the SAME bf16 tcgen05 kernels as
    one conv over 3 x the input channels — rows [hi | lo | hi] against weights [Whi; Whi; Wlo] — with fp32 accumulation in
    TMEM.  Measured per-operator error vs the fp32 oracle ~1e-5 (bar 1e-4): the tensor-core parity mode, ~3 x the conv flops
    of 'bf16' instead of the FFMA kernels' 13 x slowdown.  dgrad the same way on dy; wgrad = three pair-list launches
    (xhi, dyhi), (xlo, dyhi), (xhi, dylo) summed."""

    @staticmethod
    def forward(ctx, feats, weight, kmap: KernelMap, transposed: bool):
        _need_cuda(feats, weight)
        feats = feats.contiguous().float()
        weight = weight.contiguous().float()
        K, cin, cout = weight.shape
        if not transposed:
            table, n_dst = kmap.nbr, kmap.n_out
        else:
            table, n_dst = kmap.nbrT, kmap.n_in
        x3, xhi, xlo = split_bf16x3(feats, True, ctx.needs_input_grad[1])
        whi, wlo = _split_weight(weight)
        w3 = torch.cat([whi, whi, wlo], dim=1)  # [K, 3 cin, cout]
        out = _conv_gather_gemm("fwd", kmap, x3, w3, False, table, n_dst, cout, MATH_BF16, side=bool(transposed), algo_c_src=cin)
        ctx.save_for_backward(xhi, xlo, whi, wlo)
        ctx.misc = (kmap, transposed, feats.shape[0])
        return out

    @staticmethod
    def backward(ctx, grad_output):
        xhi, xlo, whi, wlo = ctx.saved_tensors
        kmap, transposed, n_src = ctx.misc
        K, cin, cout = whi.shape
        g = grad_output.contiguous().float()
        g3, ghi, glo = split_bf16x3(g, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        grad_feats = grad_weight = None
        if ctx.needs_input_grad[0]:
            bwd_table = kmap.nbr if transposed else kmap.nbrT
            wd3 = torch.cat([whi, whi, wlo], dim=2)  # [K, cin, 3 cout], read transposed by the kernel: rows of 3 cout -> cin
            grad_feats = _conv_gather_gemm("dgrad", kmap, g3, wd3, True, bwd_table, n_src, cin, MATH_BF16, side=not transposed,
                                           algo_c_src=cout)
        if ctx.needs_input_grad[1]:
            flat = kmap.flat_pairs
            parts = []
            dk, dflag = kmap.dense_hint()
            for j, (xa, gb) in enumerate(((xhi, ghi), (xlo, ghi), (xhi, glo))):
                dw = torch.empty((K, cin, cout), dtype=torch.float32, device=g.device)
                _timed("wgrad", kmap, g.shape[0], K, cin if j == 0 else 0, cout, lambda: check(lib().u2_conv_wgrad_pairs_dense(
                    xa.data_ptr(), cin, gb.data_ptr(), cout, kmap.nbr.data_ptr(), kmap.nbr.shape[1], kmap.n_out, K,
                    flat.data_ptr(), kmap.nbsizes.data_ptr(), int(transposed), dw.data_ptr(), MATH_BF16, dk, _ptr(dflag), _st())))
                parts.append(dw)
            grad_weight = parts[1].add_(parts[2]).add_(parts[0])  # small terms first
        return grad_feats, grad_weight, None, None


def sparse_conv(feats, weight, kmap: KernelMap, transposed: bool = False, math: Optional[int] = None):
    math = _state["math"] if math is None else math
    if math == MATH_BF16X3:
        K, cin, cout = weight.shape
        if feats.is_cuda and bf16x3_supported(cin, cout, K):
            return ConvolutionX3Fn.apply(feats, weight, kmap, transposed)
        math = MATH_FP32  # shapes the bf16 tiles cannot hold (the Cin = 4 stem conv): the FFMA kernels, fp32-exact
    return ConvolutionFn.apply(feats, weight, kmap, transposed, math)


# -------------------------------------------------------------------------------- SyncBatchNorm statistics exchange
class _PeerExchange:
    """All-reduce of the short fp64 statistics vectors of SyncBatchNorm over NVLink peer memory (csrc/syncbn.cu): torch
    symmetric memory provides the mapped buffers, one single-block kernel per exchange does the rest (no NCCL launch)."""

    def __init__(self, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(lib().u2_syncbn_buffer_bytes(), dtype=torch.uint8, device=dev)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.world, self.rank = int(self.handle.world_size), int(self.handle.rank)
        if self.world > lib().u2_syncbn_max_world():
            raise RuntimeError("group larger than the exchange buffer layout")
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's flags are zero before anybody publishes
        self.ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.seq = 0

    def all_reduce_(self, vals: torch.Tensor) -> None:
        assert vals.dtype == torch.float64 and vals.is_contiguous() and vals.numel() <= lib().u2_syncbn_max_len()
        self.seq += 1
        check(lib().u2_syncbn_exchange(vals.data_ptr(), vals.numel(), self.ptrs, self.world, self.rank, self.seq, _st()))
        _count()


_exchanges = {}


def stats_all_reduce_(vals: torch.Tensor, group) -> None:
    """Sum `vals` (fp64, short) over the ranks of `group`, in place: peer-memory exchange when the ranks share an NVLink
    domain and torch symmetric memory is available (U2_SYNCBN_TRANSPORT=auto|peer|nccl), else one NCCL all-reduce."""
    import os
    import torch.distributed as dist
    key = id(group)
    ex = _exchanges.get(key, False)
    if ex is False:
        mode = os.environ.get("U2_SYNCBN_TRANSPORT", "auto")
        ex = None
        if mode != "nccl" and vals.is_cuda and dist.get_backend(group) == "nccl":
            try:
                ex = _PeerExchange(group)
            except Exception as e:  # no symmetric memory on this system / group: the NCCL collective does the job
                if mode == "peer":
                    raise
                ex = None
                import warnings
                warnings.warn(f"SyncBatchNorm peer-memory exchange unavailable ({e!r}); using NCCL all_reduce")
        _exchanges[key] = ex
    if ex is not None:
        ex.all_reduce_(vals)
    else:
        dist.all_reduce(vals, group=group)


# -------------------------------------------------------------------------------- batch norm (+ReLU)
_pending_counters = []


def _count_batch(bn) -> None:
    """num_batches_tracked += 1.  Modules marked by fusion.optimize defer it to the end of the model's forward,
    where all counters are bumped by one multi-tensor kernel instead of one tiny kernel per layer."""
    if not bn.track_running_stats or bn.num_batches_tracked is None:
        return
    if getattr(bn, "_u2_lazy_counter", False) and len(_pending_counters) < 4096:
        _pending_counters.append(bn.num_batches_tracked)
    else:
        bn.num_batches_tracked.add_(1)


def flush_bn_counters() -> None:
    if _pending_counters:
        torch._foreach_add_(_pending_counters, 1)
        _count()
        _pending_counters.clear()


def _bn_scratch(c: int, device) -> torch.Tensor:
    """Per-CTA partial sums of the two BatchNorm reductions (consumed by the fold kernel of the same call)."""
    return _ws("bn", lib().u2_bn_scratch_bytes(c), device)


class BatchNormFn(Function):
    """Training-mode BatchNorm over [n, C] features with an optional fused ReLU and an optional
    process group (SyncBatchNorm semantics: statistics over the rows of all ranks)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu, group):
        _need_cuda(x, gamma, beta)
        x = x.contiguous().float()
        n, c = x.shape
        st = _st()
        sums = torch.empty(2 * c + 1, dtype=torch.float64, device=x.device)
        scratch = _bn_scratch(c, x.device)
        check(lib().u2_bn_stats(x.data_ptr(), n, c, sums.data_ptr(), scratch.data_ptr(), scratch.numel(), st))
        if group is not None:
            stats_all_reduce_(sums, group)
        y = torch.empty_like(x)
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        invstd = torch.empty(c, dtype=torch.float32, device=x.device)
        check(lib().u2_bn_apply(x.data_ptr(), n, c, sums.data_ptr(), float(eps), float(momentum), gamma.data_ptr(),
                                beta.data_ptr(), int(relu), y.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                _ptr(running_mean), _ptr(running_var), _st()))
        _count(3)
        ctx.save_for_backward(x, gamma, beta, mean, invstd, sums)
        ctx.misc = (relu, group)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, invstd, sums = ctx.saved_tensors
        relu, group = ctx.misc
        dy = dy.contiguous().float()
        n, c = x.shape
        st = _st()
        dsum = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        scratch = _bn_scratch(c, x.device)
        check(lib().u2_bn_bwd_reduce(dy.data_ptr(), x.data_ptr(), n, c, mean.data_ptr(), invstd.data_ptr(),
                                     gamma.data_ptr(), beta.data_ptr(), int(relu), None, dsum.data_ptr(),
                                     scratch.data_ptr(), scratch.numel(), st))
        dparam = dsum.float()   # local sums: DDP averages parameter grads
        dbeta, dgamma = dparam[:c], dparam[c:]
        if group is not None:
            dsum = dsum.clone()
            stats_all_reduce_(dsum, group)
        dx = torch.empty_like(x)
        check(lib().u2_bn_bwd_apply(dy.data_ptr(), x.data_ptr(), n, c, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), dsum.data_ptr(), sums.data_ptr() + 16 * c, int(relu), dx.data_ptr(),
                                    _st()))
        _count(2)
        return dx, dgamma, dbeta, None, None, None, None, None, None


# ------------------------------------------------------------------ conv -> BatchNorm(+ReLU), one autograd node
def stash_bf16(x: torch.Tensor, xb: torch.Tensor) -> None:
    """Remember the bf16 copy a kernel produced together with `x`; `bf16_view` hands it to the next conv
    as long as `x` has not been modified in place since."""
    x._u2_bf16 = (xb, x._version)


def bf16_view(x: torch.Tensor) -> torch.Tensor:
    st = getattr(x, "_u2_bf16", None)
    if st is not None and st[1] == x._version and st[0].shape == x.shape:
        return st[0]
    return cast_bf16(x)


class ConvBNReLUFn(Function):
    """Sequential(Conv3d, BatchNorm[, ReLU]) of core/models/build_blocks.py:21-84 as one node (bf16 math):
    forward : weights re-tiled once for both directions -> conv (epilogue leaves the column sums of Y) -> [all-reduce]
              -> normalise(+residual)(+ReLU), fp32 + bf16 out
    backward: BN reduce -> [all-reduce] -> BN dx written as bf16 only -> dgrad (+ the gradient of the input's other
              consumer, added in the conv epilogue) + wgrad
    i.e. no statistics pass over Y, no cast pass before the next conv or before dgrad/wgrad, the fp32 gradient of Y never
    exists, and no separate accumulation pass for an input that feeds this conv and a shortcut.
    want_alias: also return the input `feats` as a third output; whatever consumes that alias (the ResidualBlock shortcut)
    sends its gradient back through THIS node, where it is added inside the dgrad kernel."""

    @staticmethod
    def forward(ctx, feats, feats_bf16, weight, gamma, beta, residual, running_mean, running_var, momentum, eps, relu,
                group, kmap: KernelMap, transposed: bool, want_alias: bool = False):
        K, cin, cout = weight.shape
        weight = weight.contiguous()
        if not transposed:
            table, n_dst, side = kmap.nbr, kmap.n_out, False
        else:
            table, n_dst, side = kmap.nbrT, kmap.n_in, True
        dev = feats_bf16.device
        n_src = feats_bf16.shape[0]
        ld = table.shape[1]
        st = _st()
        l = lib()
        y = torch.empty((n_dst, cout), dtype=torch.float32, device=dev)
        # both weight blobs now (W[k] for this conv, W[k]^T for its dgrad): the parameter does not change before backward
        sbytes = l.u2_conv_scratch_bytes(n_dst, K, cin, cout, MATH_BF16)
        need_dgrad = ctx.needs_input_grad[0]
        arena = weight_blobs.get(weight) if isinstance(weight, torch.nn.Parameter) else None
        if arena is not None:
            blobs, blob_dgrad = arena
        else:
            blobs = torch.empty(2 * sbytes if need_dgrad else sbytes, dtype=torch.uint8, device=dev)
            check(l.u2_conv_pretile(weight.data_ptr(), K, cin, cout, MATH_BF16, blobs.data_ptr(),
                                    blobs.data_ptr() + sbytes if need_dgrad else None, st))
            blob_dgrad = blobs[sbytes:] if need_dgrad else None
            _count(2 if need_dgrad else 1)
        if _state["sort_tiles"]:
            tab, perm, _ = kmap.sorted_tables(side)
            rows = ld
        else:
            tab, perm, rows = table, None, n_dst
        parts = l.u2_conv_tile_stats_parts(rows)
        tstats = _ws("tstats", parts * 2 * cout * 4, dev)
        _timed("fwd", kmap, n_dst, K, cin, cout, lambda: check(l.u2_conv_fwd_stats(
            feats_bf16.data_ptr(), n_src, cin, None, 0, tab.data_ptr(), _ptr(perm), ld, n_dst, K, cout,
            y.data_ptr(), MATH_BF16, blobs.data_ptr(), sbytes, tstats.data_ptr(), tstats.numel(), st)))
        sums = torch.empty(2 * cout + 1, dtype=torch.float64, device=dev)
        check(l.u2_bn_stats_from_tiles(tstats.data_ptr(), parts, cout, n_dst, sums.data_ptr(), st))
        if group is not None:
            stats_all_reduce_(sums, group)
        z = torch.empty_like(y)
        zb = torch.empty((n_dst, cout), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((2, cout), dtype=torch.float32, device=dev)  # saved mean, invstd
        if residual is not None:
            residual = residual.contiguous()
            assert residual.shape == y.shape and residual.dtype == torch.float32, (residual.shape, y.shape)
        check(l.u2_bn_apply_dual(y.data_ptr(), n_dst, cout, sums.data_ptr(), float(eps), float(momentum),
                                 gamma.data_ptr(), beta.data_ptr(), int(relu), _ptr(residual), z.data_ptr(),
                                 zb.data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * cout, _ptr(running_mean),
                                 _ptr(running_var), st))
        _count(4)
        # with a residual the ReLU mask cannot be recomputed from y alone: keep the output z for it
        ctx.save_for_backward(feats_bf16, weight, y, gamma, beta, stats, sums,
                              z if (residual is not None and relu) else None,
                              blob_dgrad if need_dgrad else None)
        ctx.misc = (kmap, transposed, relu, group, residual is not None)
        ctx.mark_non_differentiable(zb)
        ctx.set_materialize_grads(False)  # no zero-filled "gradient" for the bf16 copy / an unused alias
        if want_alias:
            return z, zb, feats
        return z, zb

    @staticmethod
    def backward(ctx, dz, _dzb, d_alias=None):
        xb, weight, y, gamma, beta, stats, sums, zmask, blob_dgrad = ctx.saved_tensors
        kmap, transposed, relu, group, has_res = ctx.misc
        n_ret = 15
        if dz is None:
            return (d_alias,) + (None,) * (n_ret - 1)
        dz = dz.contiguous().float()
        K, cin, cout = weight.shape
        n, c = y.shape
        dev = y.device
        l = lib()
        mean_p, invstd_p = stats.data_ptr(), stats.data_ptr() + 4 * c
        dsum = torch.empty(2 * c, dtype=torch.float64, device=dev)
        scratch = _bn_scratch(c, dev)
        check(l.u2_bn_bwd_reduce(dz.data_ptr(), y.data_ptr(), n, c, mean_p, invstd_p,
                                 gamma.data_ptr(), beta.data_ptr(), int(relu), _ptr(zmask), dsum.data_ptr(),
                                 scratch.data_ptr(), scratch.numel(), _st()))
        dparam = dsum.float()
        dbeta, dgamma = dparam[:c], dparam[c:]
        if group is not None:
            dsum = dsum.clone()
            stats_all_reduce_(dsum, group)
        dyb = torch.empty((n, c), dtype=torch.bfloat16, device=dev)
        dres = None
        if has_res and ctx.needs_input_grad[5]:
            # without a ReLU the residual's gradient is dz itself; with one, the masked dz written by the kernel
            dres = torch.empty_like(dz) if relu else dz
        check(l.u2_bn_bwd_apply_dual(dz.data_ptr(), y.data_ptr(), n, c, mean_p, invstd_p,
                                     gamma.data_ptr(), beta.data_ptr(), dsum.data_ptr(), sums.data_ptr() + 16 * c,
                                     int(relu), _ptr(zmask), dres.data_ptr() if (dres is not None and relu) else None,
                                     None, dyb.data_ptr(), _st()))
        _count(3)
        grad_feats = grad_weight = None
        # dgrad and wgrad only share their inputs.  On the coarse strides neither fills the 148 SMs (a few hundred CTAs),
        # so there wgrad runs on a side stream next to dgrad; the main stream waits for it before the gradients leave.
        side = None
        if (ctx.needs_input_grad[0] and ctx.needs_input_grad[2]
                and 0 < _state.get("overlap_rows", 0) and n <= _state["overlap_rows"]):
            side = _side_stream(dev)
        tag = "+" if side is not None else ""  # conv timer: "+" = ran concurrently, the event time is not a solo duration
        if ctx.needs_input_grad[2]:
            grad_weight = torch.empty_like(weight)
            flat = kmap.flat_pairs
            dk, dflag = kmap.dense_hint()

            def wgrad():
                _timed("wgrad" + tag, kmap, n, K, cin, cout, lambda: check(l.u2_conv_wgrad_pairs_dense(
                    xb.data_ptr(), cin, dyb.data_ptr(), cout, kmap.nbr.data_ptr(), kmap.nbr.shape[1], kmap.n_out, K,
                    flat.data_ptr(), kmap.nbsizes.data_ptr(), int(transposed), grad_weight.data_ptr(), MATH_BF16, dk, _ptr(dflag),
                    _st())))

            if side is not None:
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    wgrad()
            else:
                wgrad()
        if ctx.needs_input_grad[0]:
            bwd_table = kmap.nbr if transposed else kmap.nbrT
            if d_alias is not None:
                d_alias = d_alias.contiguous().float()
            grad_feats = _conv_gather_gemm("dgrad" + tag, kmap, dyb, weight, True, bwd_table, xb.shape[0], cin, MATH_BF16,
                                           side=not transposed, blob=blob_dgrad, yadd=d_alias)
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)
        return (grad_feats, None, grad_weight, dgamma, dbeta, dres) + (None,) * (n_ret - 6)


_side_streams = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=key)
    return _side_streams[key]


def set_overlap_rows(rows: int) -> None:
    """dgrad / wgrad of a fused conv layer run concurrently (two streams) when the layer has at most `rows` output
    rows; 0 disables (default: every layer; measured 38.2 -> 37.0 ms per step)."""
    _state["overlap_rows"] = int(rows)


def _bn_group(bn):
    """Process group whose ranks share the statistics (SyncBatchNorm in training), else None."""
    import torch.distributed as dist
    if isinstance(bn, torch.nn.SyncBatchNorm) and bn.training and dist.is_available() and dist.is_initialized():
        pg = bn.process_group if bn.process_group is not None else dist.group.WORLD
        if dist.get_world_size(pg) > 1:
            return pg
    return None


def sparse_conv_bn_relu(feats, weight, kmap: KernelMap, transposed: bool, bn, relu: bool, residual=None, want_alias=False):
    """conv3d -> bn [-> + residual] (-> relu) on feature matrices. Fused node where the bf16 tcgen05 kernels cover
    the layer, otherwise the separate operators (same results up to summation order).
    want_alias: returns (out, alias) where alias is `feats` routed through the fused node — hand it to the input's other
    consumer (the ResidualBlock shortcut) and its gradient is added inside this conv's dgrad kernel."""
    group = _bn_group(bn)
    K, cin, cout = weight.shape
    n_dst = kmap.n_in if transposed else kmap.n_out
    fused = (_state["math"] == MATH_BF16 and _state.get("fuse_conv_bn", True) and bn.training and bn.affine
             and bn.momentum is not None
             and feats.is_cuda and feats.dtype == torch.float32 and weight.dtype == torch.float32 and n_dst > 1
             and feats.shape[0] > 0 and cout % 32 == 0 and lib().u2_bn_supported(cout)
             and lib().u2_conv_tc_shape_supported(cin, cout, K, MATH_BF16)
             and lib().u2_conv_tc_shape_supported(cout, cin, K, MATH_BF16)
             and lib().u2_conv_wgrad_pairs_supported(cin, cout, K, MATH_BF16)
             and (cout // ((cout + 255) // 256)) % 32 == 0)
    if not fused:
        if residual is None:
            out = batch_norm_relu(sparse_conv(feats, weight, kmap, transposed), bn, relu, group)
        else:
            out = batch_norm_relu(sparse_conv(feats, weight, kmap, transposed), bn, False, group) + residual
            out = torch.relu_(out) if relu else out
        return (out, feats) if want_alias else out
    momentum = float(bn.momentum)
    _count_batch(bn)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    feats = feats.contiguous()
    fb = bf16_view(feats)
    use_alias = bool(want_alias and feats.requires_grad and _state.get("fuse_grad_add", True))
    res = ConvBNReLUFn.apply(feats, fb, weight, bn.weight, bn.bias, residual, rm, rv, momentum, bn.eps,
                             relu, group, kmap, transposed, use_alias)
    stash_bf16(res[0], res[1])
    if use_alias:
        stash_bf16(res[2], fb)
        return res[0], res[2]
    return (res[0], feats) if want_alias else res[0]


def _bn_momentum(bn) -> float:
    """exponential_average_factor of torch's _BatchNorm.forward: `momentum`, or the cumulative average
    1 / num_batches_tracked when momentum is None (the counter is then bumped eagerly: one host read per call)."""
    if bn.momentum is not None:
        return float(bn.momentum)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        return 1.0 / float(int(bn.num_batches_tracked) + 1)
    return 0.0


def batch_norm_relu(x, bn: torch.nn.modules.batchnorm._BatchNorm, relu: bool = False, group=None):
    """BatchNorm(+ReLU) of a feature matrix [n, C] with the module's parameters / running statistics.
    Training mode on supported shapes runs the fused kernels (with `group`: statistics over all ranks, one fp64
    all-reduce per pass; a rank with zero rows contributes zero sums).  Shapes the kernels do not cover (C not a
    multiple of 4, C > 1024, no affine) go to torch: torch's own SyncBatchNorm when `group` is set — never local
    statistics, and every rank takes the same branch because the decision depends on the module and C only —
    else F.batch_norm."""
    c = x.shape[1]
    if bn.training and x.shape[0] == 1 and group is None:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {x.size()}")
    if bn.training and x.is_cuda and bn.affine and lib().u2_bn_supported(c) and (x.shape[0] > 0 or group is not None):
        momentum = _bn_momentum(bn)
        if bn.momentum is None and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        else:
            _count_batch(bn)
        rm = bn.running_mean if bn.track_running_stats else None
        rv = bn.running_var if bn.track_running_stats else None
        y = BatchNormFn.apply(x.float(), bn.weight, bn.bias, rm, rv, momentum, bn.eps, relu, group)
        return y if x.dtype == torch.float32 else y.to(x.dtype)  # autocast input: fp32 arithmetic, caller's dtype back
    if group is not None:
        y = torch.nn.SyncBatchNorm.forward(bn, x)  # torch's synchronized path (handles its own counter / momentum)
        return torch.relu(y) if relu else y
    factor = _bn_momentum(bn)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    use_batch_stats = bn.training or (bn.running_mean is None and bn.running_var is None)
    y = torch.nn.functional.batch_norm(x, bn.running_mean if (not bn.training or bn.track_running_stats) else None,
                                       bn.running_var if (not bn.training or bn.track_running_stats) else None,
                                       bn.weight, bn.bias, use_batch_stats, factor, bn.eps)
    return torch.relu(y) if relu else y
