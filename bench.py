#!/usr/bin/env python
"""Headline benchmark: SPVCNN cr=2.0 forward+backward(+SGD step) scans/s on synthetic
multisweep nuScenes-shape scans (BASELINE.json configs[1]; configs[2] for --gpus > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--math fp32|tf32|bf16|bf16x3]
    python bench.py --impl reference ...      # the reference-style CPU path (oracle) on host cores

One step = one training pass of the hot path over one batch (2 scans / GPU): H2D-resident
inputs -> initial voxelise -> 49 sparse convs (kernel maps rebuilt: every step sees a new
scan) -> point<->voxel transforms -> CE loss -> backward -> SGD.  One JSON line on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOAD = "nusc5_cr2.0_b2"
METRIC = "spvcnn_fwd_bwd_scans_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=None, choices=[None, "fp32", "tf32", "bf16", "bf16x3"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--pool", type=int, default=4, help="distinct pre-generated batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sync-bn", action="store_true")
    ap.add_argument("--bucket-mb", type=int, default=25, help="DDP gradient bucket size")
    ap.add_argument("--grad-bf16", action="store_true",
                    help="DDP communication hook: gradients cross NVLink as bf16 (halves the all-reduce bytes; off = the reference's fp32)")
    ap.add_argument("--no-fusion", action="store_true", help="keep torch BatchNorm/ReLU modules unfused")
    ap.add_argument("--overlap-rows", type=int, default=None,
                    help="run wgrad next to dgrad on a side stream for layers up to this many rows (default: all; 0 = off)")
    ap.add_argument("--no-fused-sgd", action="store_true", help="torch's default (foreach) SGD instead of fused=True")
    ap.add_argument("--no-residual-fusion", action="store_true", help="ResidualBlock tail as separate add / ReLU passes")
    ap.add_argument("--no-conv-bn", action="store_true", help="BatchNorm as its own node after each conv (A/B of the conv+BN fusion)")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="coordinate work (voxel keys, kernel maps, tile sorts) of each batch inside its own step instead of on the prefetch stream under the previous step's backward")
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid (ncu launch lists): no allocator pre-pass, no e2e region, no CPU baseline")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows)}


class ConvTimer:
    """CUDA-event pairs around every conv kernel launch inside the timed region."""

    def __init__(self):
        self.items = []

    def record(self, kind, kmap, n_dst, K, c_src, c_dst, e0, e1):
        # keep only the [K] pair-count tensor: holding the kernel map itself would pin ~100 MB of
        # neighbour tables per step and push the caching allocator into cudaMalloc every step
        self.items.append((kind, kmap.nbsizes, n_dst, K, c_src, c_dst, e0, e1))

    def by_shape(self, steps):
        """per (kind, K, c_src, c_dst): launches per step, average rows, ms per launch, ms per step, TFLOP/s over real pairs."""
        out, pairs_cache = {}, {}
        for kind, nbsizes, n_dst, K, cs, cd, e0, e1 in self.items:
            key = id(nbsizes)
            if key not in pairs_cache:
                pairs_cache[key] = int(nbsizes.sum().item())
            d = out.setdefault((kind, K, cs, cd), [0, 0.0, 0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
            d[2] += n_dst
            d[3] += 2.0 * pairs_cache[key] * cs * cd
        rows = sorted(out.items(), key=lambda kv: -kv[1][1])
        return [f"{k[0]:6s} K={k[1]:2d} {k[2]:4d}->{k[3]:4d} x{v[0] / steps:5.1f}/step rows {v[2] // v[0]:7d}  {v[1] / v[0]:7.4f} ms/launch "
                f"{v[1] / steps:7.3f} ms/step {v[3] / v[1] / 1e9:7.1f} TFLOP/s" for k, v in rows]

    def summary(self):
        """per kind: launches, total ms, algorithmic flops (2*M*Cs*Cd over REAL pairs only)."""
        out, pairs_cache = {}, {}
        for kind, nbsizes, n_dst, K, cs, cd, e0, e1 in self.items:
            key = id(nbsizes)
            if key not in pairs_cache:
                pairs_cache[key] = int(nbsizes.sum().item())
            d = out.setdefault(kind, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += 2.0 * pairs_cache[key] * cs * cd
        return out


def make_pool(args, w, rank, n):
    from u2mkd_b200 import scans
    pool = []
    for i in range(n):
        seeds = [1000 * rank + 10 * i + b for b in range(w["batch"])]
        c, f = scans.make_batch(seeds, w["kind"], w["sweeps"], w["voxel_size"])
        tgt = np.random.default_rng(seeds[0]).integers(0, 17, size=c.shape[0])
        pool.append((torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory(),
                     torch.from_numpy(tgt).pin_memory()))
    return pool


# ------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_step_fn(w, cr):
    """Reference-style CPU path: the oracle (C/OpenMP gather/scatter + torch.mm per offset,
    hash-map queries) driving the same SPVCNN, all host threads."""
    from oracle import ts_oracle
    from u2mkd_b200 import models
    ts_oracle.build()
    torch.set_num_threads(os.cpu_count())
    fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
    torch.manual_seed(0)
    net = fam.SPVCNN(cr=cr, pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)

    def step(c, f, t):
        x = ts_oracle.SparseTensor(f, c)
        out = net({"lidar": x})["x_vox"]
        loss = torch.nn.functional.cross_entropy(out, t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def crop_scan(c, f, t, frac):
    """Bounded sample: the `frac` of a scan's voxels nearest to the sensor (contiguous region)."""
    if frac >= 1.0:
        return c, f, t
    d = f[:, 0] ** 2 + f[:, 1] ** 2
    keep = torch.argsort(d)[: max(64, int(frac * c.shape[0]))].sort().values
    return c[keep].contiguous(), f[keep].contiguous(), t[keep].contiguous()


def one_scan(w, seed):
    from u2mkd_b200 import scans
    c, f = scans.make_batch([seed], w["kind"], w["sweeps"], w["voxel_size"])
    t = np.random.default_rng(seed).integers(0, 17, size=c.shape[0])
    return torch.from_numpy(c), torch.from_numpy(f), torch.from_numpy(t)


def run_reference(args, w):
    """Reference arm: the reference-style CPU path (oracle port) on the SAME fixed workload as our arm — whole scans,
    w["batch"] scans per step, the seeds of our arm's first pool batch — on all host threads.  No crop, no
    extrapolation: a step is a step (about 6 s on 16 cores for nusc5_cr2.0_b2)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from u2mkd_b200 import scans
    step = cpu_step_fn(w, w["cr"])
    seeds = [b for b in range(w["batch"])]  # == make_pool(rank 0, batch 0)
    c, f = scans.make_batch(seeds, w["kind"], w["sweeps"], w["voxel_size"])
    t = torch.from_numpy(np.random.default_rng(seeds[0]).integers(0, 17, size=c.shape[0]))
    c, f = torch.from_numpy(c), torch.from_numpy(f)
    for _ in range(args.warmup):
        step(c, f, t)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(c, f, t)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = w["batch"] / dt
    desc = (f"{w['batch']} whole scans per step ({c.shape[0]} voxels), SPVCNN cr={w['cr']} fwd+bwd+SGD, fp32, "
            f"oracle C/OpenMP + torch.mm on {os.cpu_count()} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload, "cr": w["cr"], "voxel_size": w["voxel_size"], "sweeps": w["sweeps"],
                       "scans_per_gpu": w["batch"], "voxels_per_step_rank0": int(c.shape[0])},
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": os.cpu_count(), "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(w):
    """~10-30 s of CPU work on rank 0: one bounded sample of the same workload through the oracle."""
    step = cpu_step_fn(w, w["cr"])
    c, f, t = one_scan(w, 7)
    cc = crop_scan(c, f, t, 0.05)
    step(*cc)
    t0 = time.perf_counter(); step(*cc); t_small = time.perf_counter() - t0
    frac = max(0.05, min(1.0, 20.0 / (t_small / 0.05)))
    sample = crop_scan(c, f, t, frac)
    t0 = time.perf_counter(); step(*sample); dt = time.perf_counter() - t0
    real_frac = sample[0].shape[0] / c.shape[0]
    return {"value": real_frac / dt, "unit": "scans/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{real_frac:.3f} of one scan ({sample[0].shape[0]} voxels), 1 fwd+bwd+SGD step, cr={w['cr']}, "
                      f"oracle C/OpenMP + torch.mm, {dt:.1f} s"}


# ------------------------------------------------------------------------------- our arm
def run_ours(args, w):
    import torch.distributed as dist
    from u2mkd_b200 import _lib, models, ops
    import u2mkd_b200.torchsparse as ts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    has_tc = bool(_lib.lib().u2_has_tensor_core_path())
    math = args.math or ("bf16" if has_tc else "fp32")
    ops.set_math(math)
    if args.overlap_rows is not None:
        ops.set_overlap_rows(args.overlap_rows)
    torch.backends.cuda.matmul.allow_tf32 = math in ("tf32", "bf16")   # fp32 / bf16x3 are the fp32-grade modes
    torch.backends.cudnn.allow_tf32 = math in ("tf32", "bf16")

    fam = models.product()
    torch.manual_seed(0)
    net = fam.SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"], num_classes=17).to(dev)
    if world > 1 and not args.no_sync_bn:
        net = fam.SparseSyncBatchNorm.convert_sync_batchnorm(net)  # train_spformer.py:79
    if not args.no_fusion:
        from u2mkd_b200 import fusion
        fusion.optimize(net, fuse_conv_bn=not args.no_conv_bn, fuse_residual=not args.no_residual_fusion)  # same module tree / parameters; BN(+ReLU) run the fused kernels
    if world > 1:
        # train_spformer.py:82-83.  With SyncBatchNorm every rank computes its running statistics from the same all-reduced
        # sums (bitwise: the exchange adds in rank order), so DDP's per-step broadcast of ~190 buffers from rank 0 is redundant
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=args.bucket_mb, broadcast_buffers=bool(args.no_sync_bn))
        if args.grad_bf16:
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    # torch's single-pass multi-tensor SGD (same update rule as train_spformer.py's optimizer, one kernel per chunk)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4,
                          fused=not args.no_fused_sgd)

    pool = make_pool(args, w, rank, args.pool)
    n_params = sum(p.numel() for p in net.parameters())

    def step(x):
        out = net({"lidar": x})["x_vox"]
        loss = torch.nn.functional.cross_entropy(out, x.targets)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {"last": 0.0}
    prefetch = not args.no_prefetch

    def timed(n_steps, resident):
        """n_steps steps; resident=True: inputs already in HBM; False: pinned host -> device inside the
        timed region plus a D2H read of the loss every step."""
        dev_pool = [tuple(a.to(dev) for a in b) for b in pool] if resident else None
        loss_host = torch.zeros(n_steps, dtype=torch.float32).pin_memory()
        barrier()
        ops.coord_prefetch.reset()   # every pass starts from the same allocator state of the prefetch stream (device is idle here)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host0 = time.perf_counter()
        def load(i):
            """Batch i as the model's input: device-resident tensors, or pinned host -> device copies on the current stream."""
            c, f, t = dev_pool[i % len(pool)] if resident else (a.to(dev, non_blocking=True) for a in pool[i % len(pool)])
            x = ts.SparseTensor(f, c)
            x.targets = t
            return x

        def begin(i):
            # coordinate-only work (and the H2D copies) of batch i on the prefetch stream, phase A: queued before step i - 1,
            # it runs as soon as step i - 2 is off the GPU; all of it is inside the timed region, once per step
            return fam.prepare_scan_begin(lambda: load(i), w["voxel_size"], w["voxel_size"])

        steplog = [] if os.environ.get("U2_BENCH_STEPLOG") else None
        prep = fam.prepare_scan_finish(begin(0)) if prefetch else None
        for i in range(n_steps):
            t0 = time.perf_counter()
            if prefetch:
                x = prep.x
                nxt = begin(i + 1) if i + 1 < n_steps else None
            else:
                x = load(i)
            t1 = time.perf_counter()
            loss = step(x)
            t2 = time.perf_counter()
            if prefetch and nxt is not None:
                prep = fam.prepare_scan_finish(nxt)   # phase B: row counts are in pinned memory by now; queue the map builders
            if steplog is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                ms_ = torch.cuda.memory_stats()
                try:
                    hs_ = torch.cuda.host_memory_stats().get("num_host_alloc", -1)
                except Exception:
                    hs_ = -1
                steplog.append((ev, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3,
                                ms_.get("num_device_alloc", -1), ms_.get("num_device_free", -1), hs_,
                                ms_.get("reserved_bytes.all.current", 0) >> 20))
            if not resident:
                # D2H of the step's result, every step, asynchronously into pinned memory (read after the
                # region's closing synchronize — a training loop logs the loss without stalling the queue)
                loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        e1.record()
        host_ms["last"] = (time.perf_counter() - t_host0) * 1e3 / max(1, n_steps)  # host time to ENQUEUE a step (no sync inside)
        barrier()
        if not resident:
            assert bool(torch.isfinite(loss_host).all()), "non-finite loss"
        if steplog:
            prev = e0
            rows = []
            for ev, a, b, c, na, nf, nh, rs in steplog:
                rows.append(f"gpu {prev.elapsed_time(ev):6.1f} | host begin {a:5.1f} step {b:5.1f} finish {c:5.1f} | cudaMalloc {na} cudaFree {nf} hostAlloc {nh} reserved {rs} MB")
                prev = ev
            print(f"[steplog resident={resident}]\n  " + "\n  ".join(rows), file=sys.stderr, flush=True)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # untimed: touch every pool batch once so that the caching allocator has seen every tensor shape
    # (a first-time shape inside the timed region would be a cudaMalloc + device synchronisation),
    # then the W warm-up steps proper
    # (with the prefetch two batches are alive at a time: 2 x pool steps show the allocator every pair, the wrap-around included;
    #  one cudaMalloc under load costs 10-140 ms — measured — so the pre-pass matters)
    if not args.quick:
        timed(2 * len(pool) if prefetch else len(pool), True)
    timed(args.warmup, True)
    sampler = ClockSampler(local)
    timer = ConvTimer()
    if rank == 0:
        sampler.start()
    # (1) the timed region proper: product defaults, no instrumentation
    ops.stats["launches"] = 0
    ms = timed(args.steps, True)
    host_enqueue_ms = host_ms["last"]
    launches = ops.stats["launches"]
    # (2) the same K steps once more with a CUDA-event pair around every conv launch (roofline / shares).  wgrad runs
    # on the main stream here: concurrent kernels would make their event times overlap instead of measuring a launch
    saved_overlap = ops._state["overlap_rows"]
    ops.set_overlap_rows(0)
    ops.conv_timer = timer
    timed(1, True)  # one untimed step in this configuration (its allocation pattern differs from pass 1)
    timer.items.clear()
    ms_instr = timed(args.steps, True)
    ops.conv_timer = None
    ops.set_overlap_rows(saved_overlap)
    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get("U2_BENCH_HOSTPROF"):
        # where the host spends its time queueing a step (untimed extra pass on EVERY rank — the step holds collectives —
        # rank 0 writes its profile; cProfile slows the host down ~2x)
        import cProfile, pstats, io
        pr = cProfile.Profile()
        pr.enable()
        timed(args.steps, True)
        pr.disable()
        if rank == 0:
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(45)
            with open(os.environ["U2_BENCH_HOSTPROF"], "w") as fh:
                fh.write(buf.getvalue())
    if args.quick:
        ms_e2e = ms
    else:
        # untimed: every pool batch once in this configuration too (fresh device tensors per step give the caching
        # allocator a different pattern; a first-time shape inside the timed region would be a cudaMalloc)
        timed(2 * len(pool) if prefetch else len(pool), False)
        ms_e2e = timed(args.steps, False)

    scans_per_step = w["batch"] * world
    value = scans_per_step * args.steps / (ms / 1e3)
    e2e = scans_per_step * args.steps / (ms_e2e / 1e3)
    h2d = int(np.mean([sum(a.numel() * a.element_size() for a in b) for b in pool]))

    if rank == 0:
        pk = peaks()
        summ = timer.summary()
        tot_ms = sum(d["ms"] for d in summ.values())
        # fwd and dgrad are the SAME kernel (conv_fwd_tc_kernel over the two neighbour tables): one class.
        # Kinds with a "+" ran concurrently with another conv kernel on a second stream (dgrad next to wgrad): their
        # event times are not solo durations, so `achieved` uses the solo launches of the class only, while the
        # class's share of the step counts every launch.
        classes = {"conv_fwd_tc_kernel (fwd+dgrad)": ("fwd", "dgrad", "dgrad+"),
                   "conv_wgrad_tc_kernel (wgrad)": ("wgrad", "wgrad+")}
        agg_all = {name: {f: sum(summ[k][f] for k in ks if k in summ) for f in ("launches", "ms", "flops")}
                   for name, ks in classes.items() if any(k in summ for k in ks)}
        agg = {name: {f: sum(summ[k][f] for k in ks if k in summ and not k.endswith("+")) for f in ("launches", "ms", "flops")}
               for name, ks in classes.items() if any(k in summ and not k.endswith("+") for k in ks)}
        dom_name = max(agg_all, key=lambda k: agg_all[k]["ms"])
        if dom_name not in agg:
            dom_name = max(agg, key=lambda k: agg[k]["ms"])
        dom = agg[dom_name]
        concurrent = sorted(k for k in summ if k.endswith("+"))
        # conv GEMMs are the only dense contraction: bound = tensor pipe; tf32 peak = 1/2 bf16 (BASELINE.md par. 2)
        # fp32 FFMA mode is reported against tf32 too; bf16x3 runs bf16 MMAs (3 per algorithmic product) -> bf16 peak
        peak_tf = (pk["bf16_sus"] if math in ("bf16", "bf16x3") else pk["bf16_sus"] / 2.0)
        achieved = dom["flops"] / (dom["ms"] / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r2_conv_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r1_conv_traffic.json")
        if os.path.exists(tpath) and math == "bf16":
            tk = json.load(open(tpath))["kernels"]
            # every instantiation of the dominant kernel family (fwd: with / without the epilogue addend), launch-weighted
            prefix = "conv_fwd_tc_kernel<128, 1" if dom_name.startswith("conv_fwd") else "conv_wgrad_tc_kernel<1"
            hits = [v for k, v in tk.items() if k.startswith(prefix)]
            if hits:
                traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in hits) / sum(v["launches"] for v in hits)
        roofline = {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                    "traffic_note": "dram__bytes_read+write per launch, ncu launch list profiles/" + os.path.basename(tpath).replace("conv_traffic.json", "launches_bf16.csv") + " (average over "
                                    "the kernel's launches of one step)" if traffic else None,
                    "flops_per_launch": dom["flops"] / dom["launches"],
                    "peak_source": f"{pk['src']} bf16 sustained" + ("" if math == "bf16" else " / 2 (tf32)"),
                    "launches_per_step": dom["launches"] / args.steps,
                    "avg_launch_ms": dom["ms"] / dom["launches"],
                    "share_of_step": agg_all[dom_name]["ms"] / ms_instr,
                    "measured_in": f"second pass over the same {args.steps} steps with a CUDA-event pair around every conv launch and "
                                   f"dgrad/wgrad on one stream: {ms_instr / args.steps:.2f} ms/step (the un-instrumented timed region: "
                                   f"{ms / args.steps:.2f} ms/step)",
                    "concurrent_kinds": concurrent,
                    "note": ("achieved / avg_launch_ms / launches_per_step: the class's SOLO launches; kinds marked '+' ran "
                             "concurrently on two streams (dgrad next to wgrad), their event times overlap and are counted "
                             "only in share_of_step / all_conv") if concurrent else None,
                    "all_conv": {k: {"ms_per_step": d["ms"] / args.steps, "tflops": d["flops"] / (d["ms"] / 1e3) / 1e12}
                                 for k, d in summ.items()},
                    "conv_share_of_step": tot_ms / ms_instr}
        if os.environ.get("U2_BENCH_LAYERS"):
            print("\n".join(timer.by_shape(args.steps)), file=sys.stderr)
        line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": math, "data": "synthetic",
                "config": {"workload": args.workload, "cr": w["cr"], "voxel_size": w["voxel_size"], "sweeps": w["sweeps"],
                           "scans_per_gpu": w["batch"], "voxels_per_step_rank0": int(np.mean([b[0].shape[0] for b in pool])),
                           "params": n_params, "optimizer": "sgd-nesterov", "loss": "cross_entropy",
                           "parallelism": f"dp{world}" + ("" if world == 1 or args.no_sync_bn else "+syncbn"),
                           "syncbn_transport": (None if world == 1 or args.no_sync_bn else
                                                ("nvlink peer memory" if any(v is not None for v in ops._exchanges.values()) else "nccl")),
                           "grad_allreduce": None if world == 1 else ("bf16" if args.grad_bf16 else "fp32"),
                           "fused_bn_relu": not args.no_fusion,
                           "fused_conv_bn": not (args.no_fusion or args.no_conv_bn),
                           "fused_residual": not (args.no_fusion or args.no_conv_bn or args.no_residual_fusion),
                           "optimizer_impl": "torch fused" if not args.no_fused_sgd else "torch foreach",
                           "dgrad_wgrad_overlap_rows": ops._state["overlap_rows"],
                           "coord_prefetch": prefetch,
                           "l2": "activations (>1 GB/step) exceed the 126 MB L2; a different scan batch every step"},
                "e2e": {"value": e2e, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from u2mkd_b200 import scans
    w = scans.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
