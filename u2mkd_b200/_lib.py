"""Loader / builder of the C-ABI library (include/u2mkd.h) and its ctypes signatures.

The library is built IN-TREE (u2mkd_b200/libu2mkd_b200.so) with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`; there is no CPU fallback: if the
shared object is missing the import of any op fails loudly.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libu2mkd_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

MATH_FP32, MATH_TF32, MATH_BF16 = 0, 1, 2


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one shared object."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    if not force and os.path.exists(SO_PATH) and all(os.path.getmtime(SO_PATH) >= os.path.getmtime(d) for d in deps):
        return SO_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    defs = ["-DU2_WITH_TC"] if os.path.exists(os.path.join(CSRC, "conv_tc.cu")) else []
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(_HERE, "build", os.path.basename(s)[:-3] + ".o")
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + defs + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-o", SO_PATH] + objs + ["-lcudart"])
    return SO_PATH


_lib = None

_i32, _i64, _sz, _p, _f = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float

_SIGNATURES = {
    "u2_last_error": (ctypes.c_char_p, []),
    "u2_version": (ctypes.c_int, []),
    "u2_has_tensor_core_path": (ctypes.c_int, []),
    "u2_hash": (ctypes.c_int, [_p, _i64, _p, _i32, _p, _p]),
    "u2_hash_table_bytes": (_sz, [_i64]),
    "u2_hash_table_build": (ctypes.c_int, [_p, _i64, _p, _sz, _p]),
    "u2_hash_table_query": (ctypes.c_int, [_p, _sz, _p, _i64, _p, _p]),
    "u2_count": (ctypes.c_int, [_p, _i64, _p, _i64, _p]),
    "u2_voxelize_fwd": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _i64, _p]),
    "u2_voxelize_bwd": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _i64, _p]),
    "u2_ti_weights": (ctypes.c_int, [_p, _p, _i64, _f, _p, _p]),
    "u2_devoxelize_fwd": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _i64, _p, _p]),
    "u2_devoxelize_bwd": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _i64, _p]),
    "u2_downsample_scratch_bytes": (_sz, [_i64]),
    "u2_downsample_coords": (ctypes.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "u2_kmap_scratch_bytes": (_sz, [_i64]),
    "u2_kmap_build": (ctypes.c_int, [_p, _i64, _p, _i64, _p, _i32, _p, _i64, _p, _i64, _p, _p, _sz, _p]),
    "u2_conv_scratch_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "u2_conv_fwd": (ctypes.c_int, [_p, _i64, _i32, _p, _i32, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _sz, _p]),
    "u2_conv_wgrad": (ctypes.c_int, [_p, _i64, _i32, _p, _i64, _i32, _p, _i64, _i32, _p, _i32, _p, _sz, _p]),
    "u2_kmap_sort_scratch_bytes": (_sz, [_i64]),
    "u2_kmap_sort_rows": (ctypes.c_int, [_p, _i64, _i64, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "u2_conv_fwd_perm": (ctypes.c_int, [_p, _i64, _i32, _p, _i32, _p, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _sz, _p]),
    "u2_coord_table_build": (ctypes.c_int, [_p, _i64, _p, _sz, _p]),
    "u2_coord_table_query": (ctypes.c_int, [_p, _sz, _p, _i64, _p, _i32, _p, _p]),
    "u2_unique_voxelize_scratch_bytes": (_sz, [_i64]),
    "u2_unique_voxelize": (ctypes.c_int, [_p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "u2_syncbn_buffer_bytes": (_sz, []),
    "u2_syncbn_max_len": (_i32, []),
    "u2_syncbn_max_world": (_i32, []),
    "u2_syncbn_exchange": (ctypes.c_int, [_p, _i32, _p, _i32, _i32, ctypes.c_uint64, _p]),
    "u2_conv_pretile": (ctypes.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p, _p]),
    "u2_conv_pretile_plan_bytes": (ctypes.c_size_t, [_i32]),
    "u2_conv_pretile_plan": (ctypes.c_int, [_i32, _p, _p, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    "u2_conv_pretile_run": (ctypes.c_int, [_p, _i32, _i64, _p]),
    "u2_conv_tc_shape_supported": (ctypes.c_int, [_i32, _i32, _i32, _i32]),
    "u2_conv_tile_stats_parts": (ctypes.c_size_t, [_i64]),
    "u2_conv_fwd_stats": (ctypes.c_int, [_p, _i64, _i32, _p, _i32, _p, _p, _i64, _i64, _i32, _i32, _p, _i32, _p, _sz, _p, _sz, _p]),
    "u2_bn_stats_from_tiles": (ctypes.c_int, [_p, _i64, _i32, _i64, _p, _p]),
    "u2_bn_apply_dual": (ctypes.c_int, [_p, _i64, _i32, _p, _f, _f, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "u2_bn_bwd_apply_dual": (ctypes.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "u2_bn_supported": (ctypes.c_int, [_i32]),
    "u2_bn_scratch_bytes": (ctypes.c_size_t, [_i32]),
    "u2_bn_stats": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _sz, _p]),
    "u2_bn_apply": (ctypes.c_int, [_p, _i64, _i32, _p, _f, _f, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "u2_bn_bwd_reduce": (ctypes.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _sz, _p]),
    "u2_bn_bwd_apply": (ctypes.c_int, [_p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "u2_cast_bf16": (ctypes.c_int, [_p, _i64, _p, _p]),
    "u2_split_bf16x3": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _p]),
    "u2_point2grid_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "u2_point2grid_fwd": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "u2_point2grid_bwd": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p]),
    "u2_pixel_gather_fwd": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p]),
    "u2_pixel_gather_bwd": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p]),
    "u2_window_attn_supported": (ctypes.c_int, [_i32, _i32]),
    "u2_window_pairs": (ctypes.c_int, [_p, _p, _i32, _p, _p, _p, _p, _p]),
    "u2_window_attn_fwd": (ctypes.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "u2_window_attn_bwd": (ctypes.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "u2_window_attn_bwd_scratch_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "u2_kmap_pairs_scratch_bytes": (_sz, [_i64]),
    "u2_kmap_pairs": (ctypes.c_int, [_p, _i64, _p, _p, _sz, _p]),
    "u2_conv_wgrad_pairs_supported": (ctypes.c_int, [_i32, _i32, _i32, _i32]),
    "u2_conv_wgrad_pairs": (ctypes.c_int, [_p, _i32, _p, _i32, _p, _i64, _i64, _i32, _p, _p, _i32, _p, _i32, _p]),
    "u2_conv_wgrad_pairs_dense": (ctypes.c_int, [_p, _i32, _p, _i32, _p, _i64, _i64, _i32, _p, _p, _i32, _p, _i32, _i32, _p, _p]),
}

EXPORTED = sorted(_SIGNATURES)


def lib():
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the CUDA extension is required; there is no CPU fallback)")
        l = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class U2Error(RuntimeError):
    pass


def check(status: int):
    if status != 0:
        raise U2Error(lib().u2_last_error().decode())
