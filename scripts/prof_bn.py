import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from u2mkd_b200 import ops
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); ts=[]
    for _ in range(reps):
        big.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for n, c in ((237144, 64), (237144, 192), (147146, 128), (65011, 256), (27338, 512)):
    x = torch.randn(n, c, device="cuda", requires_grad=True); bn = torch.nn.BatchNorm1d(c).cuda(); g = torch.randn(n, c, device="cuda")
    gb = n * c * 4 / 1e6
    f = t(lambda: ops.batch_norm_relu(x, bn, relu=True))
    y = ops.batch_norm_relu(x, bn, relu=True)
    b = t(lambda: torch.autograd.grad(y, x, g, retain_graph=True))
    bn2 = torch.nn.BatchNorm1d(c).cuda()
    f2 = t(lambda: torch.relu(bn2(x)))
    y2 = torch.relu(bn2(x)); b2 = t(lambda: torch.autograd.grad(y2, x, g, retain_graph=True))
    print(f"n={n} C={c}: fused fwd {f:.3f} ms ({3*gb/f:.0f} GB/s) bwd {b:.3f} ms ({5*gb/b:.0f} GB/s) | torch fwd {f2:.3f} bwd {b2:.3f}")
