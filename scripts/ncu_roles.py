#!/usr/bin/env python
"""Split an ncu source page of a warp-specialised kernel into roles at the USETMAXREG markers; per role: executed warp
instructions, stall samples by reason, top stalled instructions."""
import csv, collections, subprocess, sys
rep = sys.argv[1]; tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
ix = h.index("Instructions Executed"); isrc = h.index("Source"); ismp = h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
region = 0
agg = collections.defaultdict(lambda: dict(inst=0, smp=0, stalls=collections.Counter(), top=[], ops=collections.Counter()))
for r in rows[2:]:
    if len(r) <= ix: continue
    s = r[isrc].strip()
    if "USETMAXREG" in s: region += 1
    a = agg[region]
    n = int(r[ix]); sm = int(r[ismp])
    a["inst"] += n; a["smp"] += sm
    for i, c in stall_cols:
        a["stalls"][c] += int(r[i] or 0)
    a["top"].append((sm, n, s))
    op = s.split()[1] if s.startswith("@") and len(s.split()) > 1 else s.split()[0] if s else ""
    a["ops"][op.split(".")[0]] += n
for reg in sorted(agg):
    a = agg[reg]
    print("== region %d: %d warp-instr (%.0f per tile), %d samples" % (reg, a["inst"], a["inst"] / tiles, a["smp"]))
    print("   stalls:", ", ".join("%s=%d" % (k[6:], v) for k, v in a["stalls"].most_common(8)))
    print("   ops:", ", ".join("%s=%.0f" % (k, v / tiles) for k, v in a["ops"].most_common(14)))
    for sm, n, s in sorted(a["top"], reverse=True)[:8]:
        print("      samples %6d execs %9d  %s" % (sm, n, s[:100]))
