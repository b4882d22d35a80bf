#!/usr/bin/env python
"""Development tool: attribute the per-instruction counters of an `ncu --page source --csv` export (SASS view) to
source lines, using `nvdisasm -g -c` line info of the same cubin.

    cuobjdump -xelf all romanimpreprocess_b200/csrc/rip_v2.o ; nvdisasm -g -c rip_v2.sm_100a.cubin > lines.txt
    python tools/sass_lines.py lines.txt gpurun_out/fused_source.csv <mangled kernel name> [source file to annotate]
"""
import collections
import csv
import re
import sys

lines_txt, src_csv, kname = sys.argv[1:4]
annot = sys.argv[4] if len(sys.argv) > 4 else None
addr2line = {}
cur = None
infun = False
for ln in open(lines_txt, errors="replace"):
    if ln.startswith(".text."):
        infun = ln.strip().rstrip(":") == ".text." + kname
        continue
    if not infun:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst = collections.Counter()
smp = collections.Counter()
stall = collections.defaultdict(collections.Counter)
tot = tots = 0
base = None
for r in rows[2:]:
    try:
        a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
        n = int(r[ix["Instructions Executed"]])
    except Exception:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    inst[key] += n
    tot += n
    s = int(r[ix["# Samples"]] or 0)
    smp[key] += s
    tots += s
    for k in ix:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = int(r[ix[k]] or 0)
            if v:
                stall[key][k[6:]] += v
src = {}
if annot:
    for i, l in enumerate(open(annot), 1):
        src[i] = l.rstrip()
print(f"total warp instructions {tot}, samples {tots}")
for key, n in inst.most_common(60):
    st = " ".join(f"{k}:{v}" for k, v in stall[key].most_common(3))
    text = src.get(key[1], "") if annot and key[0] == annot.split("/")[-1] else ""
    print(f"{key[0]:>18}:{key[1]:<4} inst {100 * n / tot:5.1f}%  smp {100 * smp[key] / max(tots, 1):5.1f}%  [{st}]  {text.strip()[:90]}")

# optional: aggregate by named line ranges of the annotated file:  RANGES="name:lo-hi,name:lo-hi"
import os
if os.environ.get("RANGES") and annot:
    base_name = annot.split("/")[-1]
    agg = collections.Counter()
    for spec in os.environ["RANGES"].split(","):
        name, r = spec.split(":")
        lo, hi = map(int, r.split("-"))
        for (f, l), n in inst.items():
            if f == base_name and lo <= l <= hi:
                agg[name] += n
    other = collections.Counter()
    for (f, l), n in inst.items():
        if f != base_name:
            other[f] += n
    print("--- by range (thread-instructions per 4096^2 pixel) ---")
    for name, n in agg.most_common():
        print(f"{name:>14} {100 * n / tot:5.1f}%  {n * 32 / 4096**2:7.1f}")
    for f, n in other.most_common(6):
        print(f"{f:>14} {100 * n / tot:5.1f}%  {n * 32 / 4096**2:7.1f}")
