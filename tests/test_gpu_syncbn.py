"""The product's synchronized-statistics path (ops.BatchNormFn / ops.ConvBNReLUFn with a process group,
reference recipe core/models/utils.py:138-141 + train_spformer.py:77-83) against single-process BatchNorm over
the concatenated rows of all ranks: outputs, input gradients, parameter gradients, running statistics.

Two ranks.  With >= 2 GPUs (gpurun --gpus 2) the ranks sit on cuda:0 / cuda:1 and talk NCCL; on a one-GPU box both
ranks share cuda:0 and the collective runs over gloo (NCCL refuses two ranks on one device) — the product code is the
same, only the transport differs."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _coords(rank, n=5000):
    rng = np.random.default_rng(100 + rank)
    c = np.unique(np.concatenate([rng.integers(0, 26, (n + 700 * rank, 3)), np.full((n + 700 * rank, 1), rank)], 1).astype(np.int32), axis=0)
    rng.shuffle(c)
    return torch.from_numpy(np.ascontiguousarray(c))


def _feats(rank, n, c):
    return torch.from_numpy(np.random.default_rng(200 + rank).standard_normal((n, c)).astype(np.float32) + 0.3 * rank)


def _build(sync: bool):
    """conv -> BN -> ReLU -> conv -> BN (fused ConvBNReLUFn nodes in bf16) and a point-branch style
    Linear -> BatchNorm1d -> ReLU (BatchNormFn)."""
    from u2mkd_b200 import fusion, models
    import u2mkd_b200.torchsparse as gts
    fam = models.product()
    torch.manual_seed(3)
    net = torch.nn.ModuleDict({
        "vox": torch.nn.Sequential(gts.nn.Conv3d(64, 96, 3), gts.nn.BatchNorm(96), gts.nn.ReLU(True),
                                   gts.nn.Conv3d(96, 64, 3), gts.nn.BatchNorm(64)),
        "pts": torch.nn.Sequential(torch.nn.Linear(64, 48), torch.nn.BatchNorm1d(48), torch.nn.ReLU(True))})
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    if sync:
        net = fam.SparseSyncBatchNorm.convert_sync_batchnorm(net)
    fusion.optimize(net)
    return net


def _run(net, coords, feats, dev):
    import u2mkd_b200
    import u2mkd_b200.torchsparse as gts
    net.to(dev)
    u2mkd_b200.set_math("bf16")
    try:
        x = feats.to(dev).requires_grad_(True)
        y = net["vox"](gts.SparseTensor(x, coords.to(dev)))
        z = net["pts"](y.F)
        from u2mkd_b200 import ops
        ops.flush_bn_counters()  # fusion.optimize hooks this onto the forward of the module it was given (never called here)
        gy = torch.from_numpy(np.random.default_rng(5).standard_normal((1, 48)).astype(np.float32)).to(dev)
        gv = torch.from_numpy(np.random.default_rng(6).standard_normal((1, 64)).astype(np.float32)).to(dev)
        # the second term keeps sum_rows(dL/dy) away from zero (behind a BatchNorm alone it cancels exactly and the
        # gradient of the last voxel BatchNorm's bias would be rounding noise)
        ((z * gy).sum() + (y.F * gv).sum() + 0.1 * y.F.square().sum()).backward()
        torch.cuda.synchronize(dev)
    finally:
        u2mkd_b200.set_math("fp32")
    res = {"y": y.F.detach().cpu(), "z": z.detach().cpu(), "dx": x.grad.cpu()}
    res.update({"g:" + k: p.grad.cpu() for k, p in net.named_parameters()})
    res.update({"b:" + k: v.detach().cpu().clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k})
    return res


def _worker(rank, world, port, backend, outdir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _build(sync=True)
    c = _coords(rank)
    res = _run(net, c, _feats(rank, c.shape[0], 64), dev)
    torch.save(res, os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


@pytest.mark.timeout(600)
def test_sync_batchnorm_group_path_matches_concatenated_batch(cuda_lib):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    port = _free_port()
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, d)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=500)
            assert p.exitcode == 0, f"rank exited with {p.exitcode}"
        got = [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(2)]
    # single process, both ranks' rows in one batch (the batch index keeps the scans apart in every kernel map)
    cs = [_coords(r) for r in range(2)]
    net = _build(sync=False)
    ref = _run(net, torch.cat(cs), torch.cat([_feats(r, cs[r].shape[0], 64) for r in range(2)]), torch.device("cuda", 0))
    n0 = cs[0].shape[0]
    # same bf16 arithmetic; the statistics are summed in a different order (per rank, then across ranks)
    for k in ("y", "z", "dx"):
        both = torch.cat([got[0][k], got[1][k]])
        assert both.shape == ref[k].shape
        assert _rel(both, ref[k]) < 2e-3, (k, _rel(both, ref[k]))
        assert _rel(got[0][k], ref[k][:n0]) < 2e-3 and _rel(got[1][k], ref[k][n0:]) < 2e-3, k
    for k in ref:
        if k == "g:pts.0.bias":
            continue  # zero in exact arithmetic (bias in front of a BatchNorm): rounding noise on both sides
        if k.startswith("g:"):
            # parameter gradients are LOCAL sums on each rank (DistributedDataParallel averages them afterwards)
            assert _rel(got[0][k] + got[1][k], ref[k]) < 3e-3, (k, _rel(got[0][k] + got[1][k], ref[k]))
        if k.startswith("b:"):
            for r in range(2):
                if "num_batches" in k:
                    assert int(got[r][k]) == int(ref[k]) == 1
                else:
                    assert _rel(got[r][k], ref[k]) < 1e-4, (k, r)
