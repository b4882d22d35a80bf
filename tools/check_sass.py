#!/usr/bin/env python
"""Guard against ptxas contracting packed f32x2 multiplies into adds (rip_v2_core.cuh: the parity contract needs every
product and sum rounded separately).  For every function in the SASS of rip_v2.o the number of FFMA2 must equal the number
of explicit `fma.rn.f32x2` in the PTX of that function plus those of the PTX functions ptxas inlined into it (called from
it in PTX, absent from SASS).  usage: check_sass.py rip_v2.ptx rip_v2.o"""
import collections
import re
import subprocess
import sys

ptx_path, obj = sys.argv[1:3]
fma = collections.Counter()
calls = collections.defaultdict(set)
cur = None
pending_call = False
for ln in open(ptx_path):
    m = re.match(r"\s*(?:\.visible\s+|\.weak\s+)*(?:\.entry|\.func)\s+(?:\([^)]*\)\s*)?([_A-Za-z0-9$]+)", ln)
    if m and not ln.rstrip().endswith(";"):
        cur = m.group(1)
    if "fma.rn.f32x2" in ln:
        fma[cur] += 1
    if pending_call and cur:  # the callee of a PTX call sits on the line after `call.uni (retval),`
        m = re.match(r"\s*([_A-Za-z$][_A-Za-z0-9$]*)\s*,", ln)
        if m:
            calls[cur].add(m.group(1))
            pending_call = False
    if re.match(r"\s*call(\.uni)?\b", ln):
        m = re.search(r"call(?:\.uni)?\s+(?:\([^)]*\)\s*,\s*)?([_A-Za-z$][_A-Za-z0-9$]*)\s*,", ln)
        if m and cur:
            calls[cur].add(m.group(1))
        else:
            pending_call = True
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
ffma2 = collections.Counter()
funcs = set()
cur = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs.add(cur)
    if "FFMA2" in ln:
        ffma2[cur] += 1


def expected(f, seen=()):
    n = fma[f]
    for c in calls.get(f, ()):
        if c not in funcs and c not in seen:  # inlined by ptxas
            n += expected(c, seen + (f,))
    return n


bad = [(f, expected(f), ffma2[f]) for f in sorted(funcs) if expected(f) != ffma2[f]]
print(f"fma.rn.f32x2 vs FFMA2 checked in {len(funcs)} SASS functions ({sum(ffma2.values())} FFMA2): "
      + ("all equal" if not bad else f"{len(bad)} MISMATCH"))
for f, e, g in bad:
    print(f"  {f}: PTX {e}, SASS {g}")
sys.exit(1 if bad else 0)
