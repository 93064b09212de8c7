#!/usr/bin/env python
"""Per-source-line hot spots of every kernel in an `ncu --set full --import-source on` report:
python tools/source_hotspots.py <report.ncu-rep> [top N] > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True, check=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rr = list(csv.reader(raw.splitlines()))
names = [dict(zip(rr[0], r)) for r in rr[2:]]
kernel = -1
hdr = None
cur_file = None
per = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        if r[1].endswith("psim.cu"):
            kernel += 1          # every kernel's listing starts with the main translation unit
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] != "" and hdr and len(r) >= len(hdr) - 2:
        try:
            ie = int(r[hdr.index("Instructions Executed")])
            sm = int(r[hdr.index("# Samples")])
            te = int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        per.setdefault(kernel, []).append((cur_file, int(r[0]), r[1].strip()[:100], ie, sm, te))
for k in sorted(per):
    rows = per[k]
    ti = sum(x[3] for x in rows) or 1
    ts = sum(x[4] for x in rows) or 1
    d = names[k] if k < len(names) else {}
    print("== %s   %s us, %s warp instructions, %.1f active lanes" % (
        d.get("Kernel Name", "kernel %d" % k), d.get("gpu__time_duration.sum", "?"), ti, sum(x[5] for x in rows) / ti))
    print("   %samples  %instr  lanes  file:line  source")
    for x in sorted(rows, key=lambda x: -x[4])[:top]:
        print("   %5.1f   %5.1f   %4.1f   %s:%d  %s" % (100 * x[4] / ts, 100 * x[3] / ti, x[5] / max(1, x[3]), x[0], x[1], x[2]))
