#!/usr/bin/env python
"""Development tool: static SASS instruction counts of one kernel per source region (nvdisasm -g -c line info).
    python tools/sass_static.py lines.txt <mangled kernel name> [ncu source csv of the SAME cubin]
With the csv, executed warp-instruction counts and stall samples are attributed as well."""
import collections, csv, re, sys
lines_txt, kname = sys.argv[1:3]
src_csv = sys.argv[3] if len(sys.argv) > 3 else None
REG = [(266, 298, "row_async"), (304, 363, "loaders"), (388, 408, "stencil9"), (410, 440, "jump_exact"), (441, 460, "thr_band"),
       (461, 511, "jump_fast_var"), (545, 609, "jump_classify"), (610, 625, "clear_pair"), (626, 688, "jump_full"),
       (689, 721, "ramp_fit_fast"), (722, 746, "phi_extrap"), (754, 800, "a1:sat+raw+refpix"), (800, 836, "a1:z"), (836, 870, "a1:legendre"),
       (871, 892, "a1:store"), (893, 921, "stage_b"), (980, 1049, "c_tail"), (1051, 1114, "stage_c"), (1190, 1222, "stage_a0"),
       (1228, 1275, "step"), (1277, 1302, "prologue"), (107, 131, "SharedDiv"), (55, 106, "packed/minmax"), (132, 140, "u16"), (232, 241, "helpers")]
def region(key):
    if key is None: return "?"
    f, l = key
    if f == "rip_v2_core.cuh":
        for a, b, n in REG:
            if a <= l <= b: return n
        return f"core:{l}"
    return f
cur = None; infun = False; ins = []
for ln in open(lines_txt, errors="replace"):
    if ln.startswith(".text."):
        infun = ln.strip().rstrip(":") == ".text." + kname; continue
    if not infun: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
def opc(s):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s); return m.group(2) if m else s
execd = None
if src_csv:
    rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if r[ix["Instructions Executed"]].isdigit()]
    assert len(body) == len(ins), (len(body), len(ins))
    bad = sum(1 for a, r in zip(ins, body) if opc(a[1]) != opc(r[ix["Source"]].strip()))
    print("opcode mismatches vs csv:", bad)
    execd = [(int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]] or 0)) for r in body]
st = collections.Counter(); ex = collections.Counter(); sm = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for i, (a, s, key) in enumerate(ins):
    r = region(key); st[r] += 1
    if execd: ex[r] += execd[i][0]; sm[r] += execd[i][1]; ops[r][opc(s)] += execd[i][0]
    else: ops[r][opc(s)] += 1
tot = sum(ex.values()) or 1; tots = sum(sm.values()) or 1
print(f"{'region':22s} {'static':>7s} {'exec%':>7s} {'samples%':>8s}  top opcodes")
for r, n in sorted(st.items(), key=lambda kv: -(ex[kv[0]] if execd else kv[1])):
    top = ", ".join(f"{o}:{c * 100 // (sum(ops[r].values()) or 1)}%" for o, c in ops[r].most_common(6))
    print(f"{r:22s} {n:7d} {ex[r] / tot * 100:7.2f} {sm[r] / tots * 100:8.2f}  {top}")
print("total static", len(ins), "executed", tot)
