"""SphereFormer teacher (SPVCNN_SPFORMER, configs/nuscenes/train/spformer.yaml: cr 1.0, window 0.6 m / [2, 2, 120] deg-deg-m,
quant 1/24 of the window, head_dim 16) fwd + bwd + SGD on synthetic 5-sweep nuScenes-shape scans, 0.1 m voxels, batch 2:
scans/s of the CUDA path (bf16 conv kernels + fused nodes, fp32 fused window attention), the time spent in the two window
attention kernels, and the same step on the CPU oracles (ts_oracle + sptr_oracle) for reference.

    python scripts/bench_spformer.py [--steps 10] [--no-cpu]     -> gpurun_out/r2_bench_spformer.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from u2mkd_b200 import fusion, models, models_spformer, ops, scans
import u2mkd_b200.torchsparse as gts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    vs = 0.1
    kw = dict(window_size=np.array([0.6] * 3), window_size_sphere=[2., 2., 120.], quant_size=np.array([0.6 / 24] * 3),
              quant_size_sphere=[2 / 24, 2 / 24, 120 / 24], window_size_scale=[2.0, 2.0], drop_path_rate=0.3, a=0.0125, pres=vs, vres=vs,
              cr=1.0, num_classes=17)
    batches = []
    for i in range(4):
        c, f = scans.make_batch([10 * i, 10 * i + 1], "nusc", 5, vs)
        t = np.random.default_rng(i).integers(0, 17, size=c.shape[0])
        batches.append((torch.from_numpy(c), torch.from_numpy(f), torch.from_numpy(t)))
    ops.set_math("bf16")
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0)
    import copy   # the constructor scales the spherical sizes in place, like the reference's
    net = models_spformer.product().SPVCNN_SPFORMER(**copy.deepcopy(kw)).cuda()
    fusion.optimize(net)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4, fused=True)
    dev = [tuple(a.cuda() for a in b) for b in batches]

    def step(c, f, t):
        out = net({"lidar": gts.SparseTensor(f, c)})["x_vox"]
        loss = torch.nn.functional.cross_entropy(out, t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for b in dev:
        step(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(*dev[i % len(dev)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    assert bool(torch.isfinite(loss))
    # time inside the window-attention kernels: torch profiler over two steps
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(2):
            step(*dev[i % len(dev)])
        torch.cuda.synchronize()
    attn_us = sum(e.device_time_total for e in prof.key_averages() if "window_attn" in e.key) / 2
    all_us = sum(e.device_time_total for e in prof.key_averages()) / 2
    out = {"model": "SPVCNN_SPFORMER cr=1.0 (spformer.yaml), 4 SphereFormer blocks, cubic + spherical windows, contextual RPE",
           "workload": "synthetic nuScenes-shape, 5 sweeps, 0.1 m voxels, batch 2", "voxels_per_step": int(np.mean([b[0].shape[0] for b in batches])),
           "params": sum(p.numel() for p in net.parameters()), "ms_per_step": round(ms, 2), "scans_per_s": round(2e3 / ms, 2),
           "window_attention_kernels_ms_per_step": round(attn_us / 1e3, 3), "all_kernels_ms_per_step": round(all_us / 1e3, 2),
           "math": "bf16 conv (tcgen05, fused BN nodes) + fp32 fused window attention"}
    if not args.no_cpu:
        from oracle import sptr_oracle, ts_oracle
        ts_oracle.build()
        torch.set_num_threads(os.cpu_count())
        fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
        torch.manual_seed(0)
        cnet = models_spformer.build_spformer_family(fam, sptr_oracle.as_sptr_module()).SPVCNN_SPFORMER(**copy.deepcopy(kw))
        copt = torch.optim.SGD(cnet.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)
        c, f, t = batches[0]
        n1 = int((c[:, 3] == 0).sum())  # one scan of the batch

        def cstep():
            o = cnet({"lidar": ts_oracle.SparseTensor(f[:n1], c[:n1])})["x_vox"]
            l = torch.nn.functional.cross_entropy(o, t[:n1])
            copt.zero_grad(set_to_none=True)
            l.backward()
            copt.step()

        cstep()
        t0 = time.perf_counter()
        cstep()
        dt = time.perf_counter() - t0
        out["cpu_port"] = {"scans_per_s": round(1.0 / dt, 3), "s_per_scan": round(dt, 2), "threads": os.cpu_count(),
                           "what": "ts_oracle (C/OpenMP + torch.mm) + sptr_oracle (torch index ops), one scan, fp32"}
        out["speedup_vs_cpu_port"] = round(out["scans_per_s"] / out["cpu_port"]["scans_per_s"], 1)
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r2_bench_spformer.json", "w"), indent=1)


if __name__ == "__main__":
    main()
