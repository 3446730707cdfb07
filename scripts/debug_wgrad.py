import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u2mkd_b200 import ops, _lib
from u2mkd_b200.torchsparse.nn.utils import get_kernel_offsets
rng = np.random.default_rng(0)
c = np.unique(np.concatenate([rng.integers(0, 12, size=(1500, 3)), np.zeros((1500, 1), int)], 1), axis=0).astype(np.int32)
c = torch.from_numpy(c).cuda()
off = get_kernel_offsets(3, 1, 1, device="cuda")
km = ops.build_kernel_map(c, c, off)
n = c.shape[0]
print("n", n, "nbsizes", km.nbsizes.tolist())
flat = km.flat_pairs
M = int(km.nbsizes.sum())
print("flat[:10]", flat[:10].tolist(), "ld", km.nbr.shape[1])
ref_flat = torch.nonzero(km.nbr.view(-1) >= 0).view(-1)
print("flat ok:", torch.equal(flat[:M].long(), ref_flat))
for cin, cout in ((32, 64), (128, 16)):
    X = torch.randn(n, cin, device="cuda"); dY = torch.randn(n, cout, device="cuda")
    W = torch.zeros(27, cin, cout, device="cuda")
    dW = torch.empty_like(W)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.u2_conv_wgrad_pairs(X.data_ptr(), cin, dY.data_ptr(), cout, km.nbr.data_ptr(), km.nbr.shape[1], km.n_out, 27,
                                     flat.data_ptr(), km.nbsizes.data_ptr(), 0, dW.data_ptr(), 1, st))
    torch.cuda.synchronize()
    dW2 = torch.empty_like(W)
    _lib.check(L.u2_conv_wgrad(X.data_ptr(), n, cin, dY.data_ptr(), n, cout, km.nbr.data_ptr(), km.nbr.shape[1], 27, dW2.data_ptr(), 0, None, 0, st))
    torch.cuda.synchronize()
    print(cin, cout, "tc absmax", dW.abs().max().item(), "ref absmax", dW2.abs().max().item(), "maxdiff", (dW - dW2).abs().max().item())
    print(" tc[13,:2,:4]", dW[13, :2, :4].tolist()); print(" rf[13,:2,:4]", dW2[13, :2, :4].tolist())
    print(" per-k rel:", [round(((dW[k]-dW2[k]).abs().max()/dW2[k].abs().max().clamp_min(1e-9)).item(), 4) for k in range(27)])
