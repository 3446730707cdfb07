import sys, torch
import os; sys.path.insert(0, os.environ.get("REPO", "/root/repo"))
from oracle import ts_oracle
from u2mkd_b200 import models, scans, torchsparse as ts
if os.environ.get("THREADS"):
    torch.set_num_threads(int(os.environ["THREADS"]))
dev = torch.device("cuda:0")
coords, feats = scans.make_batch([0], "nusc", 1, 0.2)
torch.manual_seed(0)
gpu_fam = models.product()
cpu_fam = models.build_family(ts_oracle.as_torchsparse_modules()["torchsparse"])
net_cpu = cpu_fam.SPVCNN(cr=0.25, pres=0.2, vres=0.2, num_classes=17)
net_gpu = gpu_fam.SPVCNN(cr=0.25, pres=0.2, vres=0.2, num_classes=17)
net_gpu.load_state_dict(net_cpu.state_dict()); net_gpu.to(dev)
net_cpu.dropout = net_gpu.dropout = torch.nn.Identity()
def step(net, st_cls, device):
    net.zero_grad(set_to_none=True)
    x = st_cls(torch.from_numpy(feats).to(device), torch.from_numpy(coords).to(device))
    out = net({"lidar": x})["x_vox"]
    out.square().mean().backward()
    return out.detach().cpu(), {n: p.grad.detach().cpu().clone() for n, p in net.named_parameters()}
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
o_cpu, g_cpu = step(net_cpu, ts_oracle.SparseTensor, "cpu")
for it in range(1):
    o, g = step(net_gpu, ts.SparseTensor, dev)
    worst = max(((rel(g[n], g_cpu[n]), n) for n in g if float(g_cpu[n].abs().max()) > 1e-9), key=lambda t: t[0])
    print(it, "logits", f"{rel(o, o_cpu):.2e}", "stem0", f"{rel(g['stem.0.kernel'], g_cpu['stem.0.kernel']):.2e}", "worst", f"{worst[0]:.2e}", worst[1])

# which parameters are off (relative to the largest gradient in the model)?
gmax = max(float(v.abs().max()) for v in g_cpu.values())
rows = sorted(((rel(g[n], g_cpu[n]), n, float(g_cpu[n].abs().max()) / gmax) for n in g), key=lambda t: -t[0])
for e, n, mag in rows[:14]:
    print(f"{e:.2e}  |g|/gmax {mag:.1e}  {n}")

# truth: the same oracle in fp64
net_t = cpu_fam.SPVCNN(cr=0.25, pres=0.2, vres=0.2, num_classes=17)
net_t.load_state_dict(net_cpu.state_dict()); net_t.double(); net_t.dropout = torch.nn.Identity()
xt = ts_oracle.SparseTensor(torch.from_numpy(feats).double(), torch.from_numpy(coords))
out_t = net_t({"lidar": xt})["x_vox"]; out_t.square().mean().backward()
g_t = {n: p.grad.detach() for n, p in net_t.named_parameters()}
gmax = max(float(v.abs().max()) for v in g_t.values())
big = [n for n in g_t if float(g_t[n].abs().max()) > 1e-6 * gmax]
print("GPU fp32 vs fp64 truth:    stem0", f"{rel(g['stem.0.kernel'], g_t['stem.0.kernel']):.2e}", "worst", f"{max(rel(g[n], g_t[n]) for n in big):.2e}")
print("oracle fp32 vs fp64 truth: stem0", f"{rel(g_cpu['stem.0.kernel'], g_t['stem.0.kernel']):.2e}", "worst", f"{max(rel(g_cpu[n], g_t[n]) for n in big):.2e}")
print("threads", torch.get_num_threads())
