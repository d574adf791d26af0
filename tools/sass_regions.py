"""Per barrier-delimited region of a kernel's SASS: instruction mix (DFMA / shared / local (spill) / other).
usage: cuobjdump -sass obj | python tools/sass_regions.py [kernel-substring]"""
import re, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ""
lines = sys.stdin.read().split("\n")
on = False
reg = {"n": 0, "DFMA": 0, "DMUL": 0, "DADD": 0, "LDS": 0, "STS": 0, "LDL": 0, "STL": 0, "MUFU": 0, "BRA": 0}
idx = 0
start = 0
def flush(why, ln):
    global reg, idx, start
    if reg["n"]:
        print(f"region {idx:3d} lines {start}-{ln} ({why}): " + " ".join(f"{k}={v}" for k, v in reg.items() if v))
    idx += 1
    start = ln
    reg = {k: 0 for k in reg}
for ln, l in enumerate(lines):
    if "Function :" in l:
        on = pat in l
        if on: print(l.strip())
        continue
    if not on: continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if not m: continue
    op = m.group(2)
    reg["n"] += 1
    for k in reg:
        if k != "n" and op.startswith(k): reg[k] += 1
    if op.startswith("BAR") or op.startswith("WARPSYNC"): flush(op, ln)
flush("end", len(lines))
