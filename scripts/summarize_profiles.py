"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py gpurun_out/r1_launches_bf16.csv [gpurun_out/r1_conv_bf16.ncu-rep ...]

* launch list CSV (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)
  -> profiles/r1_launches_bf16_by_kernel.md (+ the per-launch average DRAM traffic of the conv kernels,
     profiles/r1_conv_traffic.json, which bench.py reports as roofline.traffic)
* .ncu-rep files (ncu --set full) -> one row of key metrics per captured launch in profiles/r1_ncu_<name>.md
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("launch__grid_size", "grid"),
]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name[:80]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    iu = hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6,
             "second": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        per[r[iid]][r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)  # -> bytes / nanoseconds
        names[r[iid]] = short(r[ik])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    tag = os.path.splitext(os.path.basename(path))[0]
    with open(os.path.join(OUT, tag + "_by_kernel.md"), "w") as f:
        f.write(f"# {tag}: per-kernel totals of one ncu launch list ({len(per)} launches, {tot / 1e6:.2f} ms of kernel time)\n\n")
        f.write("ncu serialises launches and times them cold-cache: compare SHARES, not absolutes.\n")
        f.write("DRAM = dram__bytes_read.sum + dram__bytes_write.sum, MB per launch (average over the launches).\n\n")
        f.write("| kernel | launches | time (ms) | share | DRAM MB / launch |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / a[0] / 1e6:.2f} |\n")
    traffic = {k: {"launches": a[0], "dram_bytes_per_launch": a[2] / a[0], "ms_total": a[1] / 1e6, "share": a[1] / tot}
               for k, a in agg.items() if k.startswith("conv_")}
    json.dump({"source": os.path.basename(path), "note": "ncu launch list of a short bench.py run",
               "kernels": traffic}, open(os.path.join(OUT, tag.split("_")[0] + "_conv_traffic.json"), "w"), indent=1)
    print("wrote", tag + "_by_kernel.md", "and", tag.split("_")[0] + "_conv_traffic.json")
    return agg


def ncu_rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print("no data in", path)
        return
    hdr, units = rows[0], rows[1]
    tag = os.path.splitext(os.path.basename(path))[0]
    with open(os.path.join(OUT, tag + ".md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, key metrics per captured launch\n\n")
        cols = [(k, s) for k, s in KEYS if k in hdr]
        f.write("| kernel | " + " | ".join(f"{s} ({units[hdr.index(k)]})" for k, s in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            f.write("| `" + short(r[hdr.index("Kernel Name")]) + "` | " + " | ".join(r[hdr.index(k)] for k, _ in cols) + " |\n")
    print("wrote", tag + ".md")


def main():
    os.makedirs(OUT, exist_ok=True)
    for p in sys.argv[1:]:
        if p.endswith(".csv"):
            launches(p)
        elif p.endswith(".ncu-rep"):
            ncu_rep(p)


if __name__ == "__main__":
    main()
