#!/usr/bin/env python
"""Turn what a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
  launches CSV of `ncu --metrics gpu__time_duration.sum ... --csv`  -> per-kernel summary CSV
  one `ncu --set full` report of the roofline kernel                 -> key metrics text + k2_traffic.json
Usage: python tools/summarize_profiles.py launches <in.csv> <out_summary.csv>
       python tools/summarize_profiles.py full <in.ncu-rep> <out.txt> [--traffic profiles/k2_traffic.json]"""
import collections
import csv
import json
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace(",", ";")
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for name, v in tot.most_common():
            f.write("%s,%d,%.1f,%.3f\n" % (name, cnt[name], v, v / total))


def full(src, dst, traffic=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, last = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, last))
    u = dict(zip(hdr, units))
    with open(dst, "w") as f:
        f.write("# key metrics of the LAST launch in %s (ncu --set full --clock-control none --import-source on)\n" % src.split("/")[-1])
        f.write("Kernel Name\t\t%s\n" % d["Kernel Name"])
        for k in hdr:
            if k in KEEP or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
                f.write("%s\t%s\t%s\n" % (k, u[k], d[k]))
    if traffic:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]]
        old = json.load(open(traffic))
        old.update({"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                    "source": "%s (one ncu --set full capture)" % dst})
        json.dump(old, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--traffic" else None)
