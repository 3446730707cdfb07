"""BASELINE.json configs[3]: sparse Conv3d layer sweep on a SemanticKITTI-shape scan (0.05 m voxels):
k=3 submanifold / k=2 stride-2 / k=2 stride-2 transposed, Cin = Cout = C in {32, 64, 128, 256}, at tensor strides 1 and 2.
fwd, dgrad and wgrad timed alone (CUDA events, L2 flushed between launches), flops = 2 * pairs * C * C (real pairs only),
against the measured bf16 peak; the oracle's CPU time for the same layer next to it (all host threads).

    python scripts/layer_sweep.py [--math bf16] [--reps 5] [--no-cpu]      -> gpurun_out/r2_layer_sweep_kitti.json / .md
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from u2mkd_b200 import ops, scans
import u2mkd_b200.torchsparse as gts


def timeit(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts_ = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    return float(np.median(ts_))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--math", default="bf16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) \
        else {"bf16_tflops": 1590.0}
    peak = pk["bf16_tflops"] / (1 if args.math == "bf16" else 2)  # kernels timed alone: the burst figure
    ops.set_math(args.math)
    w = scans.WORKLOADS["kitti1"]
    coords, _ = scans.make_batch([0], w["kind"], w["sweeps"], w["voxel_size"])
    c1 = torch.from_numpy(coords).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if not args.no_cpu:
        from oracle import ts_oracle
        ts_oracle.build()
        torch.set_num_threads(os.cpu_count())
    rows = []
    for ts_stride in (1, 2):
        if ts_stride == 1:
            cs = c1
        else:
            q = c1.clone()
            q[:, :3] = q[:, :3] // ts_stride * ts_stride
            cs = ops.downsample_coords(q, (1, 1, 1))
        n = cs.shape[0]
        for kind, ks, st, tr in (("k3 sub", 3, 1, False), ("k2 s2", 2, 2, False), ("k2 s2 transposed", 2, 2, True)):
            for C in (32, 64, 128, 256):
                torch.manual_seed(0)
                x = gts.SparseTensor(torch.randn(n, C, device="cuda"), cs, ts_stride)
                x.cmaps[x.s] = x.C
                down = gts.nn.Conv3d(C, C, ks, st).cuda()
                if tr:
                    mid = down(x)  # builds the stride-2 map the transposed conv reuses
                    conv = gts.nn.Conv3d(C, C, ks, st, transposed=True).cuda()
                    inp = gts.SparseTensor(torch.randn(mid.F.shape[0], C, device="cuda"), mid.C, mid.s)
                    inp.cmaps, inp.kmaps = mid.cmaps, mid.kmaps
                else:
                    conv, inp = down, x
                feats = inp.F.detach().requires_grad_(True)
                inp.F = feats
                out = conv(inp)
                key = [k for k in inp.kmaps if k[1] == (ks,) * 3][0]
                kmap = inp.kmaps[key]
                pairs = int(kmap.nbsizes.sum())
                g = torch.randn_like(out.F)
                t_fwd = timeit(lambda: conv(inp), args.reps, flush)
                # backward pieces alone: autograd.grad on one input at a time
                t_dgrad = timeit(lambda: torch.autograd.grad(conv(inp).F, feats, g), args.reps, flush) - t_fwd
                t_wgrad = timeit(lambda: torch.autograd.grad(conv(inp).F, conv.kernel, g), args.reps, flush) - t_fwd
                fl = 2.0 * pairs * C * C
                r = {"tensor_stride": ts_stride, "layer": kind, "C": C, "n_in": int(inp.F.shape[0]), "n_out": int(out.F.shape[0]),
                     "pairs": pairs, "fwd_ms": round(t_fwd, 4), "dgrad_ms": round(t_dgrad, 4), "wgrad_ms": round(t_wgrad, 4),
                     "fwd_tflops": round(fl / t_fwd / 1e9, 1), "fwd_frac_of_peak": round(fl / t_fwd / 1e9 / peak, 3)}
                if not args.no_cpu:
                  try:
                    co = ts_oracle.Conv3d(C, C, ks, st, transposed=tr)
                    xo = ts_oracle.SparseTensor(inp.F.detach().cpu(), inp.C.cpu(), inp.s)
                    if tr:  # the oracle needs its own stride-2 map
                        x0 = ts_oracle.SparseTensor(x.F.detach().cpu(), x.C.cpu(), x.s)
                        x0.cmaps[x0.s] = x0.C
                        m0 = ts_oracle.Conv3d(C, C, ks, st)(x0)
                        xo = ts_oracle.SparseTensor(inp.F.detach().cpu(), m0.C, m0.s)
                        xo.cmaps, xo.kmaps = m0.cmaps, m0.kmaps
                    with torch.no_grad():
                        co(xo)
                        t0 = time.perf_counter()
                        co(xo)
                        r["cpu_fwd_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
                    r["cpu_threads"] = os.cpu_count()
                  except Exception as e:  # the CPU column is informative only
                    r["cpu_error"] = repr(e)[:200]
                rows.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"math": args.math, "peak_tflops": peak, "rows": rows}, open("gpurun_out/r2_layer_sweep_kitti.json", "w"), indent=1)
    with open("gpurun_out/r2_layer_sweep_kitti.md", "w") as f:
        f.write(f"# r2_layer_sweep_kitti: BASELINE configs[3], SemanticKITTI-shape scan ({c1.shape[0]} voxels at 0.05 m), math={args.math}\n\n"
                f"Each kernel timed alone (CUDA events, L2 flushed); flops = 2 x real pairs x C x C; peak = {peak:.0f} TFLOP/s (measured burst). "
                "dgrad / wgrad = (fwd + that backward piece) - fwd.  cpu = the oracle's forward (C/OpenMP gather/scatter + torch.mm).\n\n"
                "| stride | layer | C | n_in | n_out | pairs | fwd ms | dgrad ms | wgrad ms | fwd TFLOP/s | of peak | cpu fwd ms |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            f.write(f"| {r['tensor_stride']} | {r['layer']} | {r['C']} | {r['n_in']} | {r['n_out']} | {r['pairs']} | {r['fwd_ms']} | {r['dgrad_ms']} | "
                    f"{r['wgrad_ms']} | {r['fwd_tflops']} | {r['fwd_frac_of_peak']} | {r.get('cpu_fwd_ms', '')} |\n")


if __name__ == "__main__":
    main()
