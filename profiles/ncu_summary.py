#!/usr/bin/env python
"""Turn one `ncu --set full` report into the summaries kept under profiles/.  A report may hold several
kernels (the two-kernel opaque stage is raster_opaque_kernel<0> + resolve_kernel, captured in one ncu pass);
per kernel <k>:
   <tag>_<k>_details.txt  (ncu --page details), <tag>_<k>_raw.json (selected raw metrics),
   <tag>_<k>_lines.txt (per-source-line shares, ncu_lines.py)
and the dram traffic of the STAGE (sum over the report's kernels) per frame -> raster_traffic.json.
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

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(m):
    v, u = float(m["value"].replace(",", "")), m["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = ncu("--page", "raw", "--csv")
rows = list(csv.reader(io.StringIO(raw)))
names, units = rows[0], rows[1]
ik = names.index("Kernel Name")
kernels = []
for vals in rows[2:]:
    if len(vals) <= ik:
        continue
    full = vals[ik].split("(")[0].split("::")[-1].replace("void ", "").strip()  # e.g. raster_opaque_kernel<1>
    short = full.replace("<", "_").replace(">", "")                              # file names: raster_opaque_kernel_1
    if short in [k for k, _, _ in kernels]:
        continue  # (one launch per kernel is summarised: the first)
    kernels.append((short, full.split("<")[0], {n: {"value": v, "unit": u} for n, u, v in zip(names, units, vals) if n in WANT}))

read = write = 0.0
parts = []
for short, base, out in kernels:
    json.dump(out, open(os.path.join(here, f"{tag}_{short}_raw.json"), "w"), indent=1)
    open(os.path.join(here, f"{tag}_{short}_details.txt"), "w").write(ncu("--page", "details", "--kernel-name", base))
    src = ncu("--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", base)
    tmp = f"/tmp/{tag}_{short}_both.csv"
    open(tmp, "w").write(src)
    lines = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), tmp, "60"], capture_output=True, text=True).stdout
    open(os.path.join(here, f"{tag}_{short}_lines.txt"), "w").write(lines)
    r, w = to_bytes(out["dram__bytes_read.sum"]), to_bytes(out["dram__bytes_write.sum"])
    read, write = read + r, write + w
    parts.append(f"{short}: {out['gpu__time_duration.sum']['value']} {out['gpu__time_duration.sum']['unit']}, "
                 f"dram read {r / 1e6:.1f} MB + write {w / 1e6:.1f} MB")
    print(short, out["gpu__time_duration.sum"], "issue active", out["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"])
    print(lines[:400])

traffic = read + write
print("dram traffic of the stage per launch:", traffic)
if key:
    p = os.path.join(here, "raster_traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[key] = {"bytes_per_frame": int(traffic / views),
              "source": (f"profiles/{tag}_*_raw.json (one ncu --set full pass over the raster stage of {views} views; "
                         + "; ".join(parts) + ")")}
    json.dump(t, open(p, "w"), indent=1)
