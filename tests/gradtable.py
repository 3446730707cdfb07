"""Per-parameter gradient table of one SPVCNN training step: product (CUDA) against the fp64 CPU oracle.

Used by tests/test_gpu_bench_parity.py and runnable on its own:

    python tests/gradtable.py [--workload nusc5_cr2.0_b2] [--math bf16|tf32|fp32] [--no-fusion] [--out gpurun_out/x.json]

Norms: per tensor  max|a - b| / max(max|b|, eps)  ("max-norm rel"), and next to it the
relative L2 error ||a - b||_2 / ||b||_2, which does not hinge on one worst element.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch


def max_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) if a.numel() else 0.0


def l2_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)) if a.numel() else 0.0


def one_scan(workload: str, seed: int):
    from u2mkd_b200 import scans
    w = scans.WORKLOADS[workload]
    c, f = scans.make_batch([seed], w["kind"], w["sweeps"], w["voxel_size"])
    t = np.random.default_rng(seed).integers(0, 17, size=c.shape[0])
    return w, torch.from_numpy(c), torch.from_numpy(f), torch.from_numpy(t)


def oracle_step(workload: str, seed: int, init_seed: int = 0, dtype=torch.float64, rounding=None):
    """The CPU oracle: (state_dict fp32, logits, {name: grad}).  dtype=float64 is the ground truth (thread-count
    independent to 1e-13).  The YARDSTICKS are the oracle's own reduced-precision runs: dtype=float32 (the reference's
    arithmetic) and `rounding` = 'bf16' / 'tf32' (fp64 oracle whose conv GEMM operands are rounded at the same points as
    the product's tensor-core modes, oracle/ts_oracle.py OPERAND_ROUNDING)."""
    from oracle import ts_oracle
    from u2mkd_b200 import models
    ts_oracle.build()
    w, c, f, t = one_scan(workload, seed)
    fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
    torch.manual_seed(init_seed)
    net = fam.SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    net.to(dtype)
    net.dropout = torch.nn.Identity()
    ts_oracle.OPERAND_ROUNDING = rounding
    try:
        out = net({"lidar": ts_oracle.SparseTensor(f.to(dtype), c)})["x_vox"]
        torch.nn.functional.cross_entropy(out, t).backward()
    finally:
        ts_oracle.OPERAND_ROUNDING = None
    return state, out.detach(), {k: p.grad.detach().clone() for k, p in net.named_parameters()}


def product_step(workload: str, seed: int, state, math: str, fused: bool = True):
    """The CUDA path as bench.py runs it (fusion.optimize, math mode), same weights, dropout off."""
    import u2mkd_b200
    import u2mkd_b200.torchsparse as ts
    from u2mkd_b200 import fusion, models
    w, c, f, t = one_scan(workload, seed)
    net = models.product().SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17)
    net.load_state_dict(state)
    net.cuda()
    net.dropout = torch.nn.Identity()
    if fused:
        fusion.optimize(net)
    u2mkd_b200.set_math(math)
    try:
        out = net({"lidar": ts.SparseTensor(f.cuda(), c.cuda())})["x_vox"]
        torch.nn.functional.cross_entropy(out, t.cuda()).backward()
        torch.cuda.synchronize()
    finally:
        u2mkd_b200.set_math("fp32")
    return out.detach().cpu(), {k: p.grad.detach().cpu() for k, p in net.named_parameters()}


YARDSTICK = {"fp32": dict(dtype=torch.float32), "tf32": dict(rounding="tf32"), "bf16": dict(rounding="bf16")}


def table(workload: str, seed: int, math: str, fused: bool = True, ref=None, yard=None):
    """Product (math mode) against the fp64 oracle, per gradient tensor; next to it (`yard_*`) how far the oracle's own
    run in the matching precision (YARDSTICK[math]) is from the same fp64 result."""
    state, out_o, grads_o = ref if ref is not None else oracle_step(workload, seed)
    out_g, grads_g = product_step(workload, seed, state, math, fused)
    _, out_y, grads_y = yard if yard is not None else oracle_step(workload, seed, **YARDSTICK[math])
    rows = [{"name": k, "shape": list(grads_o[k].shape), "max_rel": max_rel(grads_g[k], grads_o[k]),
             "l2_rel": l2_rel(grads_g[k], grads_o[k]), "yard_max_rel": max_rel(grads_y[k], grads_o[k]),
             "yard_l2_rel": l2_rel(grads_y[k], grads_o[k]), "ref_absmax": float(grads_o[k].abs().max())} for k in grads_o]
    return {"workload": workload, "seed": seed, "math": math, "fused": fused, "voxels": int(out_o.shape[0]),
            "logits_max_rel": max_rel(out_g, out_o), "logits_l2_rel": l2_rel(out_g, out_o),
            "yard_logits_max_rel": max_rel(out_y, out_o), "params": rows}


def noise_level(tab, floor=1e-9):
    """Names of gradient tensors that are zero in exact arithmetic (the bias of a Linear in front of a BatchNorm: the
    BatchNorm backward sums to zero over the rows): both sides hold rounding noise only, a relative error is meaningless."""
    scale = float(np.median([r["ref_absmax"] for r in tab["params"]]))
    return [r["name"] for r in tab["params"] if r["ref_absmax"] < floor * scale]


def worst(tab, key="max_rel", n=5):
    return sorted(tab["params"], key=lambda r: -r[key])[:n]


def describe(tab) -> str:
    skip = set(noise_level(tab))
    live = [r for r in tab["params"] if r["name"] not in skip]
    med = lambda k: float(np.median([r[k] for r in live]))
    lines = [f"{tab['workload']} seed {tab['seed']} math={tab['math']} fused={tab['fused']} voxels={tab['voxels']}: "
             f"logits max-rel {tab['logits_max_rel']:.2e} (oracle in this precision: {tab['yard_logits_max_rel']:.2e}); "
             f"{len(live)} gradient tensors (+{len(skip)} zero-in-exact-arithmetic), max-rel median {med('max_rel'):.2e} "
             f"(oracle in this precision: {med('yard_max_rel'):.2e}), l2-rel median {med('l2_rel'):.2e} ({med('yard_l2_rel'):.2e})"]
    for r in sorted(live, key=lambda r: -r["max_rel"] / max(r["yard_max_rel"], 1e-30))[:5]:
        lines.append(f"   worst vs yardstick: max-rel {r['max_rel']:.2e} (oracle {r['yard_max_rel']:.2e}), l2 {r['l2_rel']:.2e} "
                     f"(oracle {r['yard_l2_rel']:.2e}), |ref|max {r['ref_absmax']:.1e}  {r['name']} {r['shape']}")
    return "\n".join(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="nusc5_cr2.0_b2")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--math", default="bf16,tf32,fp32")
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--out", default="gpurun_out/gradtable.json")
    args = ap.parse_args()
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    ref = oracle_step(args.workload, args.seed)
    res = []
    for m in args.math.split(","):
        tab = table(args.workload, args.seed, m, fused=not args.no_fusion and m == "bf16", ref=ref)
        print(describe(tab), flush=True)
        res.append(tab)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
