#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` files into the per-kernel summary kept under profiles/:
   python tools/kernel_table.py out.csv traffic.json raw1.csv [raw2.csv ...]
One row per distinct kernel name and file (the mean over its captured launches); traffic.json gets
dram__bytes_read.sum + dram__bytes_write.sum per launch in bytes, keyed by the kernel's base name (first file wins)."""
import csv
import json
import re
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main(out_csv, traffic_json, raws):
    rows_out, traffic = [], {}
    for path in raws:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        groups = {}
        for r in rows[2:]:
            groups.setdefault(r[idx["Kernel Name"]], []).append(r)
        for name, rs in groups.items():
            vals = []
            for c in COLS:
                if c not in idx:
                    vals.append("")
                    continue
                xs = [float(r[idx[c]].replace(",", "")) for r in rs if r[idx[c]] not in ("", "n/a")]
                vals.append(f"{sum(xs) / len(xs):.6g}" if xs else "")
            rows_out.append([path.split("/")[-1], name, len(rs)] + vals)
            base = re.sub(r"^void ", "", name).split("<")[0].split("(")[0]
            if base not in traffic and "dram__bytes_read.sum" in idx:
                b = 0.0
                for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    xs = [float(r[idx[c]].replace(",", "")) for r in rs]
                    b += sum(xs) / len(xs) * UNIT.get(units[idx[c]], 1.0)
                traffic[base] = int(round(b))
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "kernel", "launches"] + COLS)
        w.writerows(rows_out)
    if traffic_json != "-":
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    for r in rows_out:
        print(r[0][:28], r[1][:60], "us", r[3], "R/W", r[4], r[5], "issue%", r[12], "lanes", r[13], "regs", r[7])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
