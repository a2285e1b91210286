#!/usr/bin/env python
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into profiles/<tag>_ncu_full_summary.csv and merge
per-kernel DRAM traffic into profiles/traffic.json (read by bench.py for roofline.traffic).
  python tools/ncu_summary.py gpurun_out/r1e_prof.ncu-rep r01e [--launches gpurun_out/r1e_launches.csv]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # the raw page, already exported on the GPU box (the report itself was too large to travel)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for d in data:
            w.writerow([d[c] for c in cols])
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tp)) if os.path.exists(tp) else {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for d in data:
        name = d[hdr.index("Kernel Name")]
        # "void <unnamed>::k_text<double, 0, 1>(...)" -> "k_text<double, 0, 1>"
        key = re.sub(r"^void ", "", name)
        key = re.sub(r"<unnamed>::|\(anonymous namespace\)::|adi::", "", key).split("(")[0]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(d[i]) * scale[units[i]]
        traffic[key] = {"dram_bytes_per_launch": tot, "grid": d[hdr.index("Grid Size")], "source": os.path.basename(out)}
        short = key.split("<")[0]
        traffic.setdefault(short, {})
        traffic[short] = {"dram_bytes_per_launch": tot, "variant": key, "source": os.path.basename(out)}
    json.dump(traffic, open(tp, "w"), indent=1, sort_keys=True)
    if "--launches" in sys.argv:
        src = sys.argv[sys.argv.index("--launches") + 1]
        keep = [l for l in open(src) if l.startswith('"') and ("adi::" in l or "k_" in l or l.startswith('"ID"'))]
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
            f.writelines(keep)
    print(out)


if __name__ == "__main__":
    main()
