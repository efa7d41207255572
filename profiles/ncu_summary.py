#!/usr/bin/env python
"""Turn one `ncu --set full` report into the summaries kept under profiles/:
   <tag>_raster_details.txt  (ncu --page details), <tag>_raster_raw.json (selected raw metrics),
   <tag>_raster_lines.txt (per-source-line shares, ncu_lines.py) and the dram traffic per launch.
usage: ncu_summary.py gpurun_out/X.ncu-rep <tag> [workload-key-for-raster_traffic.json views-in-the-captured-launch]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else None
views = int(sys.argv[4]) if len(sys.argv) > 4 else 64
here = os.path.dirname(os.path.abspath(__file__))

details = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(here, f"{tag}_raster_details.txt"), "w").write(details)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size"]
out = {}
for n, u, v in zip(names, units, vals):
    if n in want:
        out[n] = {"value": v, "unit": u}
json.dump(out, open(os.path.join(here, f"{tag}_raster_raw.json"), "w"), indent=1)


def to_bytes(m):
    v, u = float(m["value"].replace(",", "")), m["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


traffic = to_bytes(out["dram__bytes_read.sum"]) + to_bytes(out["dram__bytes_write.sum"])
print("dram traffic per launch:", traffic)
if key:
    p = os.path.join(here, "raster_traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[key] = {"bytes_per_frame": int(traffic / views),
              "source": (f"profiles/{tag}_raster_raw.json (ncu --set full, one raster launch of {views} views: "
                         f"dram read {to_bytes(out['dram__bytes_read.sum']) / 1e6:.1f} MB + write "
                         f"{to_bytes(out['dram__bytes_write.sum']) / 1e6:.1f} MB)")}
    json.dump(t, open(p, "w"), indent=1)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp = f"/tmp/{tag}_both.csv"
open(tmp, "w").write(src)
lines = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), tmp, "60"], capture_output=True, text=True).stdout
open(os.path.join(here, f"{tag}_raster_lines.txt"), "w").write(lines)
print(lines[:600])
