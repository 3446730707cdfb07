"""torch.profiler breakdown of one SPVCNN training step (GPU kernel time by kernel + CPU wall)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from u2mkd_b200 import models, ops, scans
import u2mkd_b200.torchsparse as ts

math = sys.argv[1] if len(sys.argv) > 1 else "bf16"
ops.set_math(math)
torch.backends.cuda.matmul.allow_tf32 = math != "fp32"
w = scans.WORKLOADS["nusc5_cr2.0_b2"]
dev = torch.device("cuda")
net = models.product().SPVCNN(cr=w["cr"], pres=w["voxel_size"], vres=w["voxel_size"]).to(dev)
from u2mkd_b200 import fusion
if "--no-fusion" not in sys.argv: fusion.optimize(net, fuse_conv_bn="--no-conv-bn" not in sys.argv)
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)
pool = []
for i in range(2):
    c, f = scans.make_batch([10 * i, 10 * i + 1], w["kind"], w["sweeps"], w["voxel_size"])
    pool.append((torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev),
                 torch.from_numpy(np.random.default_rng(i).integers(0, 17, size=c.shape[0])).to(dev)))

def step(c, f, t):
    out = net({"lidar": ts.SparseTensor(f, c)})["x_vox"]
    loss = torch.nn.functional.cross_entropy(out, t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()

for i in range(3):
    step(*pool[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4):
    step(*pool[i % 2])
t_cpu = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"4 steps: cpu-side issue {t_cpu*250:.1f} ms/step, wall {t_all*250:.1f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        step(*pool[i % 2])
    torch.cuda.synchronize()
from torch.autograd import DeviceType
ev = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA and e.self_device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print(f"total kernel time {tot/2e3:.2f} ms/step over {sum(e.count for e in ev)//2} launches")
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:90]:
    print(f"{e.self_device_time_total/2e3:9.3f} ms/step  x{e.count//2:4d}  {e.key[:110]}")

# GPU busy time (union of kernel intervals) vs wall: how long does the GPU sit idle inside a step?
try:
    iv = sorted((e.time_range.start, e.time_range.end) for e in prof.events() if e.device_type == DeviceType.CUDA)
    busy, cur_s, cur_e = 0.0, None, None
    for a, b in iv:
        if cur_e is None or a > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    if cur_e is not None:
        busy += cur_e - cur_s
    span = iv[-1][1] - iv[0][0]
    gaps = []
    cur_e = None
    for a, b in iv:
        if cur_e is not None and a > cur_e:
            gaps.append(a - cur_e)
        cur_e = b if cur_e is None else max(cur_e, b)
    big = sorted(gaps, reverse=True)[:12]
    print(f"GPU busy {busy/2e3:.2f} ms/step of a {span/2e3:.2f} ms/step span: idle {(span-busy)/2e3:.2f} ms/step in {len(gaps)//2} gaps/step; "
          f"largest gaps (us): {[round(g, 1) for g in big]}")
except Exception as e:
    print("busy/idle analysis failed:", repr(e))
