#!/usr/bin/env python
"""Summarise ncu output into the markdown tables kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv>            # --metrics gpu__time_duration.sum pass
    python tools/ncu_summary.py full <report.ncu-rep> [regex]       # --set full capture (needs `ncu` on PATH)
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1e-3)
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[i_name])[:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"Total {tot / 1e3:.2f} ms over {n} launches.\n")
    print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


def full(path, pattern=None):
    if path.endswith(".csv"):  # `ncu -i report.ncu-rep --page raw --csv` done on the GPU box (the reports are too big to travel)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if pattern and not re.search(pattern, name):
            continue
        print(f"### `{re.sub(r'[(].*', '', name)[:80]}`\n")
        print("| metric | value |\n|---|---:|")
        for m, label in FULL_METRICS:
            if m in ix:
                print(f"| {label} (`{m}`) | {r[ix[m]]} {units[ix[m]]} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
