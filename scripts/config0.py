"""BASELINE.json configs[0]: SPVCNN cr=0.5 FORWARD on one synthetic nuScenes-shape scan (one sweep, 34,720 rays, 0.1 m
voxels) — the reference's own CPU-runnable case (BASELINE.md par. 3 item 5).  GPU: this library (eval-free training-mode
forward, the same module tree), every math mode; CPU: the oracle port of the torchsparse v1.4.0 CPU path (C/OpenMP
gather/scatter + hash-map queries + torch.mm per offset) on all host threads.  Also fwd+bwd for both.

    python scripts/config0.py      -> gpurun_out/r2_config0.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from u2mkd_b200 import fusion, models, ops, scans
import u2mkd_b200.torchsparse as gts


def main():
    w = scans.WORKLOADS["nusc1_cr0.5"]
    coords, feats = scans.make_batch([0], w["kind"], w["sweeps"], w["voxel_size"])
    out = {"config": "BASELINE.json configs[0]", "workload": "nusc1_cr0.5", "voxels": int(coords.shape[0]), "cr": w["cr"],
           "voxel_size": w["voxel_size"], "gpu": {}, "cpu": {}}
    c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
    for math in ("bf16", "tf32", "bf16x3", "fp32"):
        ops.set_math(math)
        torch.backends.cuda.matmul.allow_tf32 = math in ("tf32", "bf16")
        torch.manual_seed(0)
        net = models.product().SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17).cuda()
        if math == "bf16":
            fusion.optimize(net)

        def fwd():
            with torch.no_grad():
                return net({"lidar": gts.SparseTensor(f, c)})["x_vox"]

        def fwd_bwd():
            net.zero_grad(set_to_none=True)
            net({"lidar": gts.SparseTensor(f, c)})["x_vox"].square().mean().backward()

        res = {}
        for name, fn in (("fwd_ms", fwd), ("fwd_bwd_ms", fwd_bwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts_ = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts_.append(e0.elapsed_time(e1))
            res[name] = round(float(np.median(ts_)), 3)
        res["scans_per_s_fwd"] = round(1e3 / res["fwd_ms"], 1)
        out["gpu"][math] = res
        print(math, res, flush=True)
    ops.set_math("fp32")
    # CPU point: the oracle on all host threads
    from oracle import ts_oracle
    ts_oracle.build()
    torch.set_num_threads(os.cpu_count())
    fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
    torch.manual_seed(0)
    net = fam.SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17)
    cc, ff = torch.from_numpy(coords), torch.from_numpy(feats)

    def cfwd():
        with torch.no_grad():
            return net({"lidar": ts_oracle.SparseTensor(ff, cc)})["x_vox"]

    def cfwd_bwd():
        net.zero_grad(set_to_none=True)
        net({"lidar": ts_oracle.SparseTensor(ff, cc)})["x_vox"].square().mean().backward()

    for name, fn in (("fwd_ms", cfwd), ("fwd_bwd_ms", cfwd_bwd)):
        for _ in range(2):
            fn()
        ts_ = []
        for _ in range(5):
            t0 = time.perf_counter()
            fn()
            ts_.append((time.perf_counter() - t0) * 1e3)
        out["cpu"][name] = round(float(np.median(ts_)), 1)
    out["cpu"]["threads"] = os.cpu_count()
    out["cpu"]["kind"] = "port (oracle: C/OpenMP + torch.mm, restated torchsparse v1.4.0 CPU path)"
    out["cpu"]["scans_per_s_fwd"] = round(1e3 / out["cpu"]["fwd_ms"], 2)
    out["speedup_fwd_bf16_vs_cpu"] = round(out["cpu"]["fwd_ms"] / out["gpu"]["bf16"]["fwd_ms"], 1)
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r2_config0.json", "w"), indent=1)


if __name__ == "__main__":
    main()
