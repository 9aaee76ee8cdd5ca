#!/usr/bin/env python3
"""Turns ncu outputs into the committed summaries under profiles/.

  launches <ncu --csv launch list> <out.csv> <command line that was profiled>
  full     <out.md> <title> <a.ncu-rep> [<b.ncu-rep> ...]     (needs `ncu` on PATH to read the reports)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(src, dst, command):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = OrderedDict()
    total = 0.0
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        ns = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = ns / 1e6 if unit in ("ns", "nsecond") else (ns / 1e3 if unit in ("us", "usecond") else ns)
        e = per.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += ms
        total += ms
    with open(dst, "w") as f:
        f.write("# ncu launch list of `%s` (C2 workload, 1 x B200)\n" % command)
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised:\n")
        f.write("# compare SHARES, not absolutes.  Captured launches: %d, total %.1f ms\n" % (len(rows), total))
        f.write("kernel,launches,total_ms,share\n")
        for name, (n, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.3f,%.3f\n" % (name, n, ms, ms / total if total else 0))


def full(dst, title, reports):
    with open(dst, "w") as f:
        f.write("# %s\n" % title)
        for rep in reports:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(out)))
            head, units = rows[0], rows[1]
            for row in rows[2:]:
                d = dict(zip(head, row))
                u = dict(zip(head, units))
                f.write("\n## %s  (%s)\n" % (d["Kernel Name"], rep.split("/")[-1]))
                for k in KEYS:
                    if k in d:
                        f.write("%s = %s %s\n" % (k, d[k], u.get(k, "")))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
    else:
        sys.exit(__doc__)
