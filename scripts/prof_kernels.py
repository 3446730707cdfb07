"""Runs each hot-path kernel a few times on a realistic multisweep scan so that ncu can capture it.
Also prints CUDA-event timings + algorithmic bytes/flops for every kernel:

    python scripts/prof_kernels.py [--reps 5] [--only conv|pv|hash] [--cr 2.0]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from u2mkd_b200 import _lib, models, ops, scans
import u2mkd_b200.torchsparse as ts
from u2mkd_b200.torchsparse.nn import functional as F
from u2mkd_b200.torchsparse.nn.utils import get_kernel_offsets


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts_ = []
    for _ in range(reps):
        big.zero_()  # flush L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    return float(np.median(ts_))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="all")
    ap.add_argument("--math", default="bf16")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm, tf32 = peaks["hbm_gbs"], peaks["bf16_tflops"] / (1 if args.math == "bf16" else 2)
    ops.set_math(args.math)
    coords, feats = scans.make_batch([0, 1], "nusc", 5, 0.05)
    c = torch.from_numpy(coords).cuda()
    n = c.shape[0]
    rows = []

    def report(name, ms, bytes_=None, flops=None, note=""):
        r = {"kernel": name, "ms": round(ms, 4)}
        if bytes_ is not None:
            r["GB/s"] = round(bytes_ / ms / 1e6, 1)
            r["hbm_frac"] = round(bytes_ / ms / 1e6 / hbm, 3)
        if flops is not None:
            r["TFLOP/s"] = round(flops / ms / 1e9, 1)
            r["tensor_frac"] = round(flops / ms / 1e9 / tf32, 3)
        r["note"] = note
        rows.append(r)
        print(json.dumps(r), flush=True)

    if args.only in ("all", "hash"):
        off27 = get_kernel_offsets(3, 1, 1, device="cuda")
        report("hash_kernel", timeit(lambda: F.sphash(c), args.reps), bytes_=24 * n, note=f"n={n}")
        h = F.sphash(c)
        q = F.sphash(c, get_kernel_offsets(2, 1, 1, device="cuda"))
        cap = _lib.lib().u2_hash_table_bytes(n)
        report("kernel_hash_kernel(K=8)", timeit(lambda: F.sphash(c, get_kernel_offsets(2, 1, 1, device="cuda")), args.reps),
               bytes_=(16 + 64) * n)
        report("table build+query (sphashquery, 8n queries)", timeit(lambda: F.sphashquery(q, h), args.reps),
               bytes_=8 * n + 2 * cap + 16 * 8 * n, note="2 kernels + memset; table L2-resident")
        report("kmap build k3 (insert+query 27 offsets)", timeit(lambda: ops.build_kernel_map(c, c, off27), args.reps),
               bytes_=(16 + 4 * 27) * n + 4 * 27 * n + 2 * cap, note="nbr + nbrT tables written")
        report("downsample_coords s2 (key+sort+unique+unpack)", timeit(lambda: ops.downsample_coords(c, (2, 2, 2)), args.reps),
               bytes_=16 * n * 2 + 8 * n * 4, note="includes host sync for the row count")

    km = ops.build_kernel_map(c, c, get_kernel_offsets(3, 1, 1, device="cuda"))
    M = int(km.nbsizes.sum())
    if args.only in ("all", "conv"):
        for cin, cout in ((64, 64), (128, 128), (256, 192), (192, 192)):
            x = torch.randn(n, cin, device="cuda")
            w = torch.randn(27, cin, cout, device="cuda") * 0.05
            g = torch.randn(n, cout, device="cuda")
            mth = ops._state["math"]
            xo, go = (ops.cast_bf16(x), ops.cast_bf16(g)) if mth == 2 else (x, g)
            fl = 2.0 * M * cin * cout
            by = (n * cin + n * cout + 27 * cin * cout) * 4 + 4 * 27 * n
            report(f"conv_fwd k3 {cin}->{cout} n={n} M={M}", timeit(lambda: ops._conv_gather_gemm("fwd", km, xo, w, False, km.nbr, n, cout, mth, side=False), args.reps),
                   bytes_=by, flops=fl)
            report(f"conv_dgrad k3 {cout}->{cin}", timeit(lambda: ops._conv_gather_gemm("dgrad", km, go, w, True, km.nbrT, n, cin, mth, side=True), args.reps),
                   bytes_=by, flops=fl)
            flat = km.flat_pairs
            dw = torch.empty_like(w)
            st = torch.cuda.current_stream().cuda_stream
            if ops._state["math"] != 0:
                report(f"conv_wgrad k3 {cin}x{cout}", timeit(lambda: _lib.check(_lib.lib().u2_conv_wgrad_pairs(
                    xo.data_ptr(), cin, go.data_ptr(), cout, km.nbr.data_ptr(), km.nbr.shape[1], n, 27, flat.data_ptr(),
                    km.nbsizes.data_ptr(), 0, dw.data_ptr(), mth, st)), args.reps),
                    bytes_=(n * cin + n * cout + 27 * cin * cout) * 4 + 8 * M, flops=fl)

    if args.only in ("all", "pv"):
        fam = models.product()
        for stride, ch in ((1, 64), (4, 256), (16, 512)):
            cs = c.clone()
            cs[:, :3] = cs[:, :3] // stride * stride
            vox = ops.downsample_coords(cs, (1, 1, 1)) if stride > 1 else c
            nv = vox.shape[0]
            z = ts.PointTensor(torch.randn(n, ch, device="cuda"), c.float())
            x = ts.SparseTensor(torch.randn(nv, ch, device="cuda"), vox, stride)
            idx = F.sphashquery(F.sphash(cs), F.sphash(vox))
            cnt = F.spcount(idx.int(), nv)
            idx32 = idx.int()
            report(f"voxelize_fwd s{stride} C={ch} n_pts={n} n_vox={nv}", timeit(lambda: F.spvoxelize(z.F, idx32, cnt), args.reps),
                   bytes_=(n * ch + nv * ch) * 4 + 4 * n)
            gv = torch.randn(nv, ch, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            gin = torch.empty(n, ch, device="cuda")
            report(f"voxelize_bwd s{stride} C={ch}", timeit(lambda: _lib.check(_lib.lib().u2_voxelize_bwd(
                gv.data_ptr(), nv, ch, idx32.data_ptr(), cnt.data_ptr(), gin.data_ptr(), n, st)), args.reps),
                bytes_=(n * ch + nv * ch) * 4 + 4 * n)
            off8 = get_kernel_offsets(2, stride, 1, device="cuda")
            iq = F.sphashquery(F.sphash(cs, off8), F.sphash(vox))
            w8 = F.calc_ti_weights(z.C, iq, scale=stride).t().contiguous()
            report(f"ti_weights s{stride}", timeit(lambda: F.calc_ti_weights(z.C, iq, scale=stride), args.reps), bytes_=n * (16 + 64 + 32))
            iq8 = iq.t().contiguous().int()
            report(f"devoxelize_fwd s{stride} C={ch}", timeit(lambda: F.spdevoxelize(x.F, iq8, w8), args.reps),
                   bytes_=(nv * ch + n * ch) * 4 + 64 * n)
            gp = torch.randn(n, ch, device="cuda")
            gf = torch.empty(nv, ch, device="cuda")
            report(f"devoxelize_bwd s{stride} C={ch}", timeit(lambda: _lib.check(_lib.lib().u2_devoxelize_bwd(
                gp.data_ptr(), n, ch, iq8.data_ptr(), w8.data_ptr(), gf.data_ptr(), nv, st)), args.reps),
                bytes_=(nv * ch + n * ch) * 4 + 64 * n, note="includes the memset of the output")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/prof_kernels.json", "w"), indent=1)


if __name__ == "__main__":
    main()
