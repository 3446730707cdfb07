import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from u2mkd_b200 import fusion, models, scans, ops
import u2mkd_b200.torchsparse as gts
def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
coords, feats = scans.make_batch([3], "nusc", 1, 0.1)
ops.set_math("bf16")
outs = []
for fused in (False, False, True, True):
    torch.manual_seed(0)
    net = models.product().SPVCNN(cr=1.0, pres=0.1, vres=0.1).cuda()
    net.dropout = torch.nn.Identity()
    fusion.optimize(net, fuse_conv_bn=fused)
    acts = {}
    hooks = []
    for name, m in net.named_modules():
        if isinstance(m, torch.nn.Sequential) or name.count(".") <= 1:
            def hk(mod, i, o, name=name):
                f = o.F if hasattr(o, "F") else o
                if torch.is_tensor(f): acts[name] = f.detach().clone()
            hooks.append(m.register_forward_hook(hk))
    out = net({"lidar": gts.SparseTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda())})["x_vox"]
    outs.append((out.detach(), acts))
print("unfused vs unfused", rel_err(outs[1][0], outs[0][0]))
print("fused vs fused", rel_err(outs[3][0], outs[2][0]))
print("fused vs unfused", rel_err(outs[2][0], outs[0][0]))
for k in outs[0][1]:
    if k in outs[2][1] and outs[0][1][k].shape == outs[2][1][k].shape:
        print(f"{k:40s} {rel_err(outs[2][1][k], outs[0][1][k]):.2e}  (noise {rel_err(outs[1][1][k], outs[0][1][k]):.2e})")
