"""N>1 host logic on CPU: world_size-2 gloo run of the data-parallel step (bench.py's recipe:
per-rank scans, DistributedDataParallel gradient all-reduce) with the model mirror driven by the
CPU oracle, checked against the single-process average of the two ranks' gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(seed=0):
    from oracle import ts_oracle
    from u2mkd_b200 import models
    fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
    torch.manual_seed(seed)
    net = fam.SPVCNN(cr=0.125, pres=0.4, vres=0.4, num_classes=5)
    net.dropout = torch.nn.Identity()
    return ts_oracle, fam, net


def _loss(ts_oracle, net, rank):
    from u2mkd_b200 import scans
    c, f = scans.make_batch([100 + rank], "nusc", 1, 0.4)
    t = torch.from_numpy(np.random.default_rng(rank).integers(0, 5, size=c.shape[0]))
    out = net({"lidar": ts_oracle.SparseTensor(torch.from_numpy(f), torch.from_numpy(c))})["x_vox"]
    return torch.nn.functional.cross_entropy(out, t)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts_oracle, fam, net = _build()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    _loss(ts_oracle, ddp, rank).backward()
    if rank == 0:
        q.put({k: p.grad.numpy().copy() for k, p in net.named_parameters()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_two_ranks_gloo_matches_mean_of_local_grads():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=500)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ts_oracle, fam, net = _build()
    grads = []
    for rank in range(2):
        net.zero_grad()
        # BatchNorm running stats differ per rank but do not enter training-mode gradients
        _loss(ts_oracle, net, rank).backward()
        grads.append({k: p.grad.clone() for k, p in net.named_parameters()})
    for k in got:
        want = (grads[0][k] + grads[1][k]) / 2
        # same arithmetic, different thread counts / reduction order through 49 BN layers
        err = float((torch.from_numpy(got[k]) - want).abs().max() / want.abs().max().clamp_min(1e-6))
        assert err < 3e-2, (k, err)


def test_sync_batchnorm_conversion_keeps_parameters():
    """SparseSyncBatchNorm.convert_sync_batchnorm (core/models/utils.py:143-220) on the mirror."""
    ts_oracle, fam, net = _build()
    keys = list(net.state_dict().keys())
    conv = fam.SparseSyncBatchNorm.convert_sync_batchnorm(net)
    assert list(conv.state_dict().keys()) == keys
    n_sparse = sum(isinstance(m, fam.SparseSyncBatchNorm) for m in conv.modules())
    n_dense = sum(type(m) is torch.nn.SyncBatchNorm for m in conv.modules())
    assert n_sparse == 49 and n_dense == 3
