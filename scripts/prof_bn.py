"""Times the four BatchNorm kernels of csrc/norm.cu one by one on the layer shapes of the bench workload and
prints their HBM roofline fraction (algorithmic bytes / time / measured peak):  python scripts/prof_bn.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u2mkd_b200 import _lib, ops

L = _lib.lib()
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
HBM = json.load(open(pk)).get("hbm_gbs", 6650.0) if os.path.exists(pk) else 6650.0
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, reps=None):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps or REPS):
        big.zero_()  # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


tot = {"stats": 0.0, "apply": 0.0, "bwd_reduce": 0.0, "bwd_apply": 0.0}
ideal = dict(tot)
shapes = ((238000, 64, 2), (238000, 192, 5), (150000, 64, 5), (150000, 192, 5), (68000, 128, 5), (68000, 256, 5),
          (28000, 256, 5), (28000, 512, 5), (9500, 512, 5))
if os.environ.get("BN_SHAPES"):
    shapes = tuple(tuple(int(v) for v in sh.split("x")) + (1,) for sh in os.environ["BN_SHAPES"].split(","))
REPS = int(os.environ.get("BN_REPS", "7"))
for n, c, mult in shapes:
    x = torch.randn(n, c, device="cuda"); dy = torch.randn(n, c, device="cuda")
    y = torch.empty_like(x); yb = torch.empty(n, c, dtype=torch.bfloat16, device="cuda")
    g = torch.rand(c, device="cuda") + 0.5; b = torch.randn(c, device="cuda") * 0.1
    mean = torch.empty(c, device="cuda"); inv = torch.empty(c, device="cuda")
    sums = torch.empty(2 * c + 1, dtype=torch.float64, device="cuda"); dsum = torch.empty(2 * c, dtype=torch.float64, device="cuda")
    st = ops._st()
    scr = ops._bn_scratch(c, 'cuda')
    P = lambda a: a.data_ptr()
    r = {}
    r["stats"] = (t(lambda: _lib.check(L.u2_bn_stats(P(x), n, c, P(sums), P(scr), scr.numel(), st))), 4)
    r["apply"] = (t(lambda: _lib.check(L.u2_bn_apply_dual(P(x), n, c, P(sums), 1e-5, 0.1, P(g), P(b), 1, None, P(y), P(yb), P(mean), P(inv), None, None, st))), 10)
    r["bwd_reduce"] = (t(lambda: _lib.check(L.u2_bn_bwd_reduce(P(dy), P(x), n, c, P(mean), P(inv), P(g), P(b), 1, None, P(dsum), P(scr), scr.numel(), st))), 8)
    r["bwd_apply"] = (t(lambda: _lib.check(L.u2_bn_bwd_apply_dual(P(dy), P(x), n, c, P(mean), P(inv), P(g), P(b), P(dsum), P(sums) + 16 * c, 1, None, None, None, P(yb), st))), 10)
    line = f"n={n:6d} C={c:3d}:"
    for k, (ms, bpe) in r.items():
        gbs = n * c * bpe / ms / 1e6
        line += f"  {k} {ms*1e3:6.1f} us {gbs:5.0f} GB/s ({gbs/HBM:4.2f})"
        tot[k] += ms * mult; ideal[k] += n * c * bpe / HBM / 1e6 * mult
    print(line, flush=True)
print("weighted totals (ms/step estimate | at HBM peak):", {k: (round(v, 3), round(ideal[k], 3)) for k, v in tot.items()})
