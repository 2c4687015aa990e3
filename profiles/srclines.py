#!/usr/bin/env python
"""Per-source-line totals of an `ncu --set full --import-source on` report (needs -lineinfo):
   python profiles/srclines.py report.ncu-rep [top_n]   -> instructions executed and stall samples per CUDA line."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and r and r[0] not in ("", "Function Name"):
        d = dict(zip(hdr[4:], r[4:]))
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]), int(d["# Samples"])))
        except (ValueError, KeyError):
            pass
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
print("--- by instructions executed")
for f, n, src, ins, smp in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{ins / ti:6.3f} inst {smp / ts:6.3f} smp  {f}:{n}  {src}")
print("--- by stall samples")
for f, n, src, ins, smp in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{ins / ti:6.3f} inst {smp / ts:6.3f} smp  {f}:{n}  {src}")
