"""Where does a conv work item's time go?  Runs conv_fwd_tc_kernel on the bench scan batch with parts of its
traffic switched off (U2_CONV_DIAG; results are invalid in those modes, only the time is read):
  0 = product, 1 = no weight-blob loads, 2 = no row gathers, 3 = neither (pipeline + MMA + epilogue only),
  4 = gathers hit 128 hot rows (LSU issue cost without L2 traffic), 5 = 4 + no weight loads

    python scripts/diag_conv.py [--reps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from u2mkd_b200 import ops, scans
from u2mkd_b200.torchsparse.nn.utils import get_kernel_offsets


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts_ = []
    for _ in range(reps):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    return float(np.median(ts_))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", default="0,1,2,3,4,5")
    ap.add_argument("--shapes", default="all")
    ap.add_argument("--stats", action="store_true", help="run the conv with the fused BatchNorm statistics epilogue")
    args = ap.parse_args()
    ops.set_math("bf16")
    coords, _ = scans.make_batch([0, 1], "nusc", 5, 0.05)
    c1 = torch.from_numpy(coords).cuda()
    maps = {}
    for stride in (1, 8):
        cs = c1.clone()
        cs[:, :3] = cs[:, :3] // stride * stride
        vox = ops.downsample_coords(cs, (1, 1, 1)) if stride > 1 else c1
        km = ops.build_kernel_map(vox, vox, get_kernel_offsets(3, stride, 1, device="cuda"))
        maps[stride] = (vox.shape[0], km, int(km.nbsizes.sum()))
    rows = []
    shapes = ((1, 64, 64), (1, 128, 128), (1, 192, 192), (1, 256, 192), (8, 256, 256), (8, 512, 512), (8, 768, 512))
    if args.shapes != "all":
        shapes = tuple(tuple(int(v) for v in sh.split("x")) for sh in args.shapes.split(","))
    for stride, cin, cout in shapes:
        n, km, M = maps[stride]
        x = ops.cast_bf16(torch.randn(n, cin, device="cuda"))
        w = torch.randn(27, cin, cout, device="cuda") * 0.05
        r = {"stride": stride, "cin": cin, "cout": cout, "n": n, "pairs": M}
        for mode in [int(m) for m in args.modes.split(",")]:
            os.environ["U2_CONV_DIAG"] = str(mode)
            r[f"ms_diag{mode}"] = round(timeit(lambda: ops._conv_gather_gemm("fwd", km, x, w, False, km.nbr, n, cout, 2, side=False), args.reps), 4)
        os.environ["U2_CONV_DIAG"] = "0"
        r["tflops"] = round(2.0 * M * cin * cout / r["ms_diag0"] / 1e9, 1)
        rows.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/diag_conv.json", "w"), indent=1)


if __name__ == "__main__":
    main()
