import sys, os, numpy as np, torch, time
sys.path.insert(0, os.getcwd())
from u2mkd_b200.sptr import functional as F
from u2mkd_b200 import sptr
torch.manual_seed(0)
rng = np.random.default_rng(0)
def run(n_pts, mean_len, h, L=47, d=16, tag="", rel_on=True):
    counts = torch.from_numpy(np.maximum(1, rng.poisson(mean_len, size=max(1, n_pts // mean_len))).astype(np.int64))
    N, M = int(counts.sum()), int((counts**2).sum())
    q, k, v = (torch.randn(N, h, d, device="cuda", requires_grad=True) for _ in range(3))
    tabs = [torch.randn(L, 3, h, d, device="cuda", requires_grad=True) for _ in range(3)]
    rel = torch.randint(0, L, (M, 3), device="cuda", dtype=torch.int32)
    wo, so_ = F.window_offsets(counts.cuda())
    for rep in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        o = F.window_attention(q, k, v, wo, so_, counts.shape[0], rel, *tabs) if rel_on else F.window_attention(q, k, v, wo, so_, counts.shape[0])
        e[1].record()
        o.sum().backward()
        e[2].record()
        torch.cuda.synchronize()
    print(f"{tag} N={N} windows={counts.shape[0]} mean={mean_len} n_max={int(counts.max())} M={M/1e6:.2f}M h={h}: fwd {e[0].elapsed_time(e[1]):.3f} ms, bwd {e[1].elapsed_time(e[2]):.3f} ms", flush=True)
run(67000, 5, 1, tag="s1 cubic ")
run(67000, 24, 1, tag="s1 sphere")
run(28000, 37, 2, tag="s2 sphere")
run(8700, 36, 4, tag="s3 sphere")
run(2700, 32, 8, tag="s4 sphere")
run(67000, 100, 1, tag="big win  ")
run(67000, 5, 1, tag="s1 cubic  no tables", rel_on=False)
run(67000, 24, 1, tag="s1 sphere no tables", rel_on=False)
run(67000, 5, 1, tag="s1 cubic (again)")
run(67000, 3, 1, tag="tiny windows")
run(67000, 8, 1, tag="windows of 8")
run(67000, 5, 1, L=9, tag="s1 cubic L=9")
