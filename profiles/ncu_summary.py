"""Summarise an `ncu --set full` report: a fixed list of metrics per captured kernel, and DRAM traffic per launch.

usage: ncu -i report.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv [traffic.json-key-suffix]
"""
import csv
import json
import re
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
           "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
           "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
           "smsp__pcsamp_warps_issue_stalled_selected"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, suffix=""):
    with open(path) as f:
        rows = list(csv.reader(ln for ln in f if ln.startswith('"')))
    names, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(names)}
    seen, traffic = set(), {}
    for r in data:
        kern = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0]
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        if units[col["gpu__time_duration.sum"]] in ("us", "usecond") and dur < 10.0:
            continue          # a launch past convergence: returns at once
        if kern in seen:
            continue
        seen.add(kern)
        print(f"kernel: {kern}")
        for m in METRICS:
            if m in col:
                print(f"  {m:75s} {r[col[m]]:>14s} {units[col[m]]}")
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * SCALE[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * SCALE[units[col["dram__bytes_write.sum"]]]
        print(f"  dram traffic (read+write) per launch: {(rd + wr) / 1e6:.1f} MB\n")
        key = kern.replace("<short, 1>", "0").replace("<int, 1>", "0")
        key = re.sub(r"<(short|int)(, 0)?>", "", key).replace("<", "").replace(">", "")
        traffic[key + suffix] = rd + wr
    print("# traffic json:", json.dumps(traffic))


if __name__ == "__main__":
    main(*sys.argv[1:])
