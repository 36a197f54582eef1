#!/usr/bin/env python
"""ncu_summary.py -- turn an `ncu --set full` report of the bench command into the small JSON bench.py quotes.

  ncu -i gpurun_out/X.ncu-rep --page raw --csv > X_raw.csv      (done here if given the .ncu-rep and ncu is on PATH)
  python tools/ncu_summary.py X.ncu-rep|X_raw.csv profiles/r2_fused_vcycle_ncu_summary.json --command "..."

Keeps, per profiled launch of k_gsrb_fused*, the duration, DRAM bytes, grid, registers, stall ratios; marks the
finest-level launches (the largest grids) and averages their DRAM traffic -> `finest_level_avg_dram_gbyte_per_launch`
(`roofline.traffic`).  `kernel_source_sha16` ties the figure to the kernel sources it was measured on (bench.py drops
the figure when the fingerprint of the build differs)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def to_gbyte(value, unit):
    v = float(str(value).replace(",", ""))
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}[unit]


def main():
    src, out = sys.argv[1], sys.argv[2]
    command = sys.argv[sys.argv.index("--command") + 1] if "--command" in sys.argv else ""
    what = sys.argv[sys.argv.index("--what") + 1] if "--what" in sys.argv else ""
    kfilter = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "k_gsrb_fused"
    if src.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(src).read()
    rows = list(csv.reader(io.StringIO(text)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    launches = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names) or kfilter not in r[col["Kernel Name"]]:
            continue
        rec = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")}
        for k in KEEP:
            if k in col:
                rec[k] = {"value": r[col[k]], "unit": units[col[k]]}
        launches.append(rec)
    # the finest-level launches are the ones that write a whole finest-level array (grids differ between the variants)
    written = [to_gbyte(l["dram__bytes_write.sum"]["value"], l["dram__bytes_write.sum"]["unit"]) for l in launches]
    wmax = max(written) if written else 0.0
    finest = [i for i, w in enumerate(written) if w >= 0.9 * wmax] if written else []
    rd = [to_gbyte(launches[i]["dram__bytes_read.sum"]["value"], launches[i]["dram__bytes_read.sum"]["unit"]) for i in finest]
    wr = [to_gbyte(launches[i]["dram__bytes_write.sum"]["value"], launches[i]["dram__bytes_write.sum"]["unit"]) for i in finest]
    import bench
    d = {"command": command, "what": what, "kernel_source_sha16": bench.kernel_fingerprint(), "finest_level_launches": finest,
         "finest_level_avg_dram_gbyte_per_launch": {"read": sum(rd) / max(len(rd), 1), "write": sum(wr) / max(len(wr), 1),
                                                    "total": (sum(rd) + sum(wr)) / max(len(rd), 1)},
         "launches": launches}
    json.dump(d, open(out, "w"), indent=1)
    print(out, "finest launches", finest, d["finest_level_avg_dram_gbyte_per_launch"])


if __name__ == "__main__":
    main()
