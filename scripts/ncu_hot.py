#!/usr/bin/env python3
"""Top stall-sampled SASS instructions of an .ncu-rep source page (needs --import-source / -lineinfo)."""
import csv, subprocess, sys, io
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
a, s, w, ex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
body = []
for i, r in enumerate(rows[2:]):
    try: body.append((float(r[w]), i, r[s].strip(), float(r[ex])))
    except Exception: pass
tot = sum(b[0] for b in body) or 1
print(f"total samples {tot:.0f}, instructions {len(body)}")
for v, i, t, e in sorted(body, reverse=True)[:top]:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  #{i:4d} exec={e:10.0f}  {t[:110]}")
