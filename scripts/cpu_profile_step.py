"""cProfile of the host side of one SPVCNN training step (where do the CPU microseconds go)."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u2mkd_b200 import models, ops, scans
import u2mkd_b200.torchsparse as ts
ops.set_math("bf16")
torch.backends.cuda.matmul.allow_tf32 = True
w = scans.WORKLOADS["nusc5_cr2.0_b2"]
dev = torch.device("cuda")
net = models.product().SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"]).to(dev)
from u2mkd_b200 import fusion
fusion.optimize(net)
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)
c, f = scans.make_batch([0, 1], w["kind"], w["sweeps"], w["voxel_size"])
c, f = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)
t = torch.from_numpy(np.random.default_rng(0).integers(0, 17, size=c.shape[0])).to(dev)
def step():
    out = net({"lidar": ts.SparseTensor(f, c)})["x_vox"]
    loss = torch.nn.functional.cross_entropy(out, t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
# phase timing
def fwd_only():
    return net({"lidar": ts.SparseTensor(f, c)})["x_vox"]
torch.cuda.synchronize(); t0 = time.perf_counter(); out = fwd_only(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"fwd: cpu issue {1e3*(t1-t0):.1f} ms, wall {1e3*(t2-t0):.1f} ms")
loss = torch.nn.functional.cross_entropy(out, t)
torch.cuda.synchronize(); t0 = time.perf_counter(); loss.backward(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"bwd: cpu issue {1e3*(t1-t0):.1f} ms, wall {1e3*(t2-t0):.1f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:12000])
s2 = io.StringIO()
pstats.Stats(pr, stream=s2).sort_stats("cumulative").print_stats(40)
print(s2.getvalue()[:9000])
