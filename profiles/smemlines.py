#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts of an `ncu --set full --import-source on` report (needs -lineinfo):
   python profiles/smemlines.py report.ncu-rep [top_n]  -> lines ranked by excessive shared wavefronts (bank conflicts)."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
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
            w, i, x = int(d["L1 Wavefronts Shared"]), int(d["L1 Wavefronts Shared Ideal"]), int(d["L1 Wavefronts Shared Excessive"])
            if w:
                lines.append((cur_file, int(r[0]), r[1].strip()[:100], w, i, x, int(d["Instructions Executed"]),
                              int(d.get("stall_short_sb", 0) or 0)))
        except (ValueError, KeyError):
            pass
tw = sum(l[3] for l in lines) or 1
print(f"shared wavefronts {tw}, ideal {sum(l[4] for l in lines)}, excessive {sum(l[5] for l in lines)}")
for f, n, src, w, i, x, ins, ssb in sorted(lines, key=lambda l: -l[5])[:top]:
    print(f"{w:9d} wavefronts {i:9d} ideal {x:9d} excess {ins:8d} inst  {f}:{n}  {src}")
