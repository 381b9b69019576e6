#!/usr/bin/env python
"""Top stall lines / opcode histogram of one kernel of an .ncu-rep:  ncu_source_top.py <rep> <kernel index>"""
import collections
import csv
import subprocess
import sys

rep, want = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
b = blocks[want]
hdr, data = b[0], [r for r in b[1:] if len(r) == len(b[0])]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("kernels in report:", len(blocks), "| total samples", tot, "| SASS lines", len(data))
op_s, op_i = collections.Counter(), collections.Counter()
for r in data:
    toks = r[src].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    op_s[op] += int(r[si])
    op_i[op] += int(r[ie])
print("by samples:", op_s.most_common(16))
print("by warp-instructions:", op_i.most_common(16), "total", sum(op_i.values()))
for r in sorted(data, key=lambda r: -int(r[si]))[:24]:
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    print(r[si], r[ie], r[src].strip()[:64], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
tot_st = collections.Counter()
for r in data:
    for i in stall_cols:
        tot_st[hdr[i][6:]] += int(r[i])
print("stall totals:", tot_st.most_common(12))
