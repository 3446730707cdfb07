import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ts_oracle as oracle
from u2mkd_b200 import models, scans
import u2mkd_b200.torchsparse as gts
cr, vs, seeds = 0.5, 0.2, [0]
coords, feats = scans.make_batch(seeds, "nusc", 1, vs)
torch.manual_seed(0)
net_o = models.build_family(oracle.as_torchsparse_modules()["torchsparse"]).SPVCNN(cr=cr, pres=vs, vres=vs)
net_g = models.product().SPVCNN(cr=cr, pres=vs, vres=vs)
net_g.load_state_dict(net_o.state_dict()); net_g.cuda()
net_o.dropout = net_g.dropout = torch.nn.Identity()
target = torch.from_numpy(np.random.default_rng(0).integers(0, 17, size=coords.shape[0]))
def step(net, st_cls, dev):
    x = st_cls(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev))
    out = net({"lidar": x})["x_vox"]
    torch.nn.functional.cross_entropy(out, target.to(dev)).backward()
    return out
og = step(net_g, gts.SparseTensor, "cuda"); oo = step(net_o, oracle.SparseTensor, "cpu")
rows = []
for (n, pg), (_, po) in zip(net_g.named_parameters(), net_o.named_parameters()):
    a, b = pg.grad.cpu().double(), po.grad.double()
    rows.append(((a-b).abs().max().item()/max(b.abs().max().item(),1e-30), n, b.abs().max().item(), (a-b).abs().max().item()))
rows.sort(reverse=True)
for r in rows[:25]: print("%.3e %-50s max|g|=%.3e maxdiff=%.3e" % r)
