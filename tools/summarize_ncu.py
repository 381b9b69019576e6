#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/summarize_ncu.py <tag>     # e.g. r01

  gpurun_out/launches.csv      (ncu --metrics gpu__time_duration.sum ... --csv)  -> profiles/<tag>_launches_by_kernel.csv
  gpurun_out/prof_gemm.ncu-rep (ncu --set full -k regex:conv_gemm_kernel)        -> profiles/<tag>_gemm_full.csv
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

lp = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("mrd::<unnamed>::", "")
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches_by_kernel.csv"), "w") as fh:
        fh.write("kernel,launches,total_us,share_of_all_launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k},{v[0]},{v[1] / 1e3:.1f},{v[1] / tot:.4f}\n")
    print("launch list:", len(data), "launches,", f"{tot / 1e6:.2f} ms")

rp = os.path.join(ROOT, "gpurun_out", "prof_gemm.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out, f"{tag}_gemm_full.csv"), "w") as fh:
        fh.write(",".join(w for w, _ in idx) + "\n")
        fh.write(",".join(rows[1][i] for _, i in idx) + "\n")  # units
        for r in rows[2:]:
            fh.write(",".join('"' + r[i].replace('"', "'")[:80] + '"' if w == "Kernel Name" else r[i] for w, i in idx) + "\n")
    print("full capture:", len(rows) - 2, "kernels")
