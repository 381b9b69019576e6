#!/usr/bin/env python
"""Per-kernel SASS opcode census of libmrd_b200.so: how many tcgen05 MMA (UTCHMMA/UTC*MMA), TMEM load (LDTM),
TMA load / store (UTMALDG / UTMASTG / UBLKCP) and legacy mma.sync (HMMA) instructions each kernel contains.

    python tools/sass_census.py > profiles/r02_sass_census.csv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-rare-disease_b200", "libmrd_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMAREDG", "HMMA", "MUFU",
       "SYNCS", "REDG", "ATOMG", "RED."]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                              stdout=subprocess.PIPE, text=True).stdout.splitlines()
    names = iter(demangle)
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names, m.group(1))
            cur = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", cur)
            cur = re.sub(r"\((int|bool|unsigned int)\)", "", cur)       # template value casts
            cur = re.sub(r"\([^()]*\)\s*$", "", cur)                   # the parameter list
            cur = cur.replace("void ", "").replace("mrd::", "").strip()
            while cur in counts:
                cur += "'"
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["TOTAL"] += 1
            for o in OPS:
                if op.startswith(o.rstrip(".")) if not o.endswith(".") else op.startswith(o):
                    counts[cur][o] += 1
    cols = ["TOTAL"] + OPS
    print("kernel," + ",".join(c.rstrip(".") for c in cols))
    tot = collections.Counter()
    for k in order:
        print(k.replace(",", ";") + "," + ",".join(str(counts[k][c]) for c in cols))
        tot.update(counts[k])
    print("ALL KERNELS," + ",".join(str(tot[c]) for c in cols))


if __name__ == "__main__":
    sys.exit(main())
