#!/usr/bin/env python
"""Per-REGION totals of an `ncu --set full --import-source on` report: python profiles/srcregions.py report.ncu-rep
Regions are (file, first line, last line, label) ranges of the current sources; edit REGIONS when the code moves."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and r and r[0] not in ("", "Function Name"):
        d = dict(zip(hdr[4:], r[4:]))
        try:
            lines.append((cur, int(r[0]), int(d["Instructions Executed"]), int(d["# Samples"])))
        except (ValueError, KeyError):
            pass
ti = sum(l[2] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
agg = {}
for f, n, i, s in lines:
    key = (f, n // int(sys.argv[2]) * int(sys.argv[2])) if len(sys.argv) > 2 else (f,)
    a = agg.setdefault(key, [0, 0])
    a[0] += i; a[1] += s
print(f"total warp instructions {ti}, samples {ts}")
for k, (i, s) in sorted(agg.items()):
    if i / ti > 0.004 or s / ts > 0.004:
        print(f"{i/ti:6.3f} inst {s/ts:6.3f} smp  {k}")
