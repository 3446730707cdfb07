import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u2mkd_b200 import ops, fusion
import u2mkd_b200
def rel(a,b): a,b=a.double().cpu(),b.double().cpu(); return float((a-b).abs().max()/b.abs().max())
torch.manual_seed(0)
n,cin,cout=6000,64,512
x=torch.randn(n,cin,device='cuda'); W=torch.randn(cout,cin,device='cuda')/8
u2mkd_b200.set_math('bf16')
km=ops.identity_kernel_map(n,'cuda')
w3=W.t().contiguous().unsqueeze(0)
# plain conv path (no BN): fwd + dgrad + wgrad
xr=x.clone().requires_grad_(True); w3r=w3.clone().requires_grad_(True)
y=ops.sparse_conv(xr,w3r,km)
g=torch.randn(n,cout,device='cuda')
y.backward(g)
print('plain: y',rel(y,x.double()@W.t().double()),'dx',rel(xr.grad,g.double()@W.double()),'dw',rel(w3r.grad[0],(x.double().t()@g.double())))
for overlap in (0, 1<<30):
    ops.set_overlap_rows(overlap)
    bn=torch.nn.BatchNorm1d(cout).cuda()
    xr=x.clone().requires_grad_(True); w3r=w3.clone().requires_grad_(True)
    z=ops.sparse_conv_bn_relu(xr,w3r,km,False,bn,True)
    z.backward(g)
    # reference in fp64
    xd=x.double().requires_grad_(True); wd=W.double().requires_grad_(True)
    bnd=torch.nn.BatchNorm1d(cout).double().cuda()
    zd=torch.relu(bnd(xd@wd.t())); zd.backward(g.double())
    print('fused overlap',overlap,': z',rel(z,zd),'dx',rel(xr.grad,xd.grad),'dw',rel(w3r.grad[0],wd.grad.t()),'dgamma',rel(bn.weight.grad,bnd.weight.grad))
