#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/summarize_ncu.py <tag>     # e.g. r01

  gpurun_out/launches.csv      (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
        -> profiles/<tag>_launches_by_kernel.csv   per kernel: launches, time, share, DRAM bytes per launch
        -> profiles/<tag>_traffic.json             DRAM traffic per launch of the dominant family (bench.py reads it)
  gpurun_out/prof_gemm.ncu-rep (ncu --set full -k regex:conv_gemm_kernel)
        -> profiles/<tag>_gemm_full.csv
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
cmd_note = sys.argv[2] if len(sys.argv) > 2 else ""
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# optional: MRD_NCU_LIST / MRD_NCU_REP = other input files under gpurun_out/, MRD_NCU_SUFFIX = output name suffix
lp = os.path.join(ROOT, "gpurun_out", os.environ.get("MRD_NCU_LIST", "launches.csv"))
sfx = os.environ.get("MRD_NCU_SUFFIX", "")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mn, mu, mv = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.defaultdict(lambda: {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    for r in data:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("mrd::<unnamed>::", "")
        try:
            v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
        except ValueError:
            continue
        a = agg[name]
        if r[mn] == "gpu__time_duration.sum":
            a["n"] += 1
            a["ns"] += v
        elif r[mn] == "dram__bytes_read.sum":
            a["rd"] += v
        elif r[mn] == "dram__bytes_write.sum":
            a["wr"] += v
    tot = sum(a["ns"] for a in agg.values())
    with open(os.path.join(out, f"{tag}{sfx}_launches_by_kernel.csv"), "w") as fh:
        fh.write("kernel,launches,total_us,share_of_all_launches,dram_read_MB_per_launch,dram_write_MB_per_launch\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            n = max(a["n"], 1)
            fh.write(f'{k},{a["n"]},{a["ns"] / 1e3:.1f},{a["ns"] / tot:.4f},{a["rd"] / n / 1e6:.3f},{a["wr"] / n / 1e6:.3f}\n')
    fam = [a for k, a in agg.items() if k.startswith("conv_gemm_kernel")]
    n = sum(a["n"] for a in fam)
    if n and not sfx:
        tr = {"kernel": "conv_gemm_kernel (all instantiations)", "launches_captured": n,
              "dram_bytes_per_launch": (sum(a["rd"] for a in fam) + sum(a["wr"] for a in fam)) / n,
              "dram_read_bytes_per_launch": sum(a["rd"] for a in fam) / n,
              "dram_write_bytes_per_launch": sum(a["wr"] for a in fam) / n,
              "share_of_all_kernel_time": sum(a["ns"] for a in fam) / tot,
              "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                        "--clock-control none, whole run of: " + cmd_note}
        with open(os.path.join(out, f"{tag}_traffic.json"), "w") as fh:
            json.dump(tr, fh, indent=1)
    print("launch list:", sum(a["n"] for a in agg.values()), "launches,", f"{tot / 1e6:.2f} ms")

rp = os.path.join(ROOT, "gpurun_out", os.environ.get("MRD_NCU_REP", "prof_gemm.ncu-rep"))
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(os.path.join(out, f"{tag}{sfx}_{'gemm' if not sfx else 'kernels'}_full.csv"), "w") as fh:
        fh.write(",".join(w for w, _ in idx) + "\n")
        fh.write(",".join(rows[1][i] for _, i in idx) + "\n")  # units
        for r in rows[2:]:
            fh.write(",".join('"' + r[i].replace('"', "'")[:80] + '"' if w == "Kernel Name" else r[i] for w, i in idx) + "\n")
    print("full capture:", len(rows) - 2, "kernels")
