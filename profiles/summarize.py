#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small text summaries kept under profiles/.

  python profiles/summarize.py launches <launches.csv>          per-kernel share of device time (gpu__time_duration.sum pass)
  python profiles/summarize.py full <report.ncu-rep> [top]      key counters of an `ncu --set full` capture + hottest source lines
"""
import collections
import csv
import io
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].split("::")[-1]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    print("kernel launches: %d, summed device time %.3f ms (%s, cold-cache, serialised by ncu: compare SHARES)" % (sum(c for c, _ in agg.values()), tot / 1e6, rows[hi + 1][mu]))
    print("%-32s %8s %14s %8s %12s" % ("kernel", "launches", "total_ms", "share", "avg_us"))
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-32s %8d %14.3f %7.1f%% %12.1f" % (n, c, v / 1e6, 100 * v / tot, v / c / 1e3))


def full(rep, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("capture: %s   kernel: %s   launches captured: %d" % (os.path.basename(rep), data[0][hdr.index("Kernel Name")].split("(")[0], len(data)))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-88s %-12s %s" % (k, units[i], "  ".join(r[i] for r in data)))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    print()
    sys.stdout.flush()
    subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), tmp, str(top)])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
