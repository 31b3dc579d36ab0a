#!/usr/bin/env python
"""Text summary of an ncu report (read here, without a GPU): per kernel the raw metrics bench.py's roofline rests on, and for
one kernel the source lines with the most stall samples.

    python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep [kernel-regex-for-the-line-table] > profiles/<name>_ncu_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
           "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "k_tile"
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("ncu --set full --clock-control none --import-source on;", rep.split("/")[-1])
    for r in rows[2:]:
        print("---")
        print("  Kernel Name =", r[idx["Kernel Name"]])
        for m in METRICS:
            if m in idx:
                print("  %s = %s %s" % (m, r[idx[m]], units[idx[m]]))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + pat]))))
    cur, h = None, None
    lines = collections.defaultdict(lambda: [0, 0, 0, 0, 0, ""])
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur, h = r[1].split("/")[-1], None
            continue
        if r[0] == "Line No":
            h = r
            continue
        if h is None or r[0] == "Function Name":
            continue
        d = dict(zip(h, r))
        try:
            ln, inst, smp = int(d["Line No"]), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)
            bar, lsb, ssb = int(d.get("stall_barrier") or 0), int(d.get("stall_long_sb") or 0), int(d.get("stall_short_sb") or 0)
        except ValueError:
            continue
        v = lines[(cur, ln)]
        v[0] += inst; v[1] += smp; v[2] += bar; v[3] += lsb; v[4] += ssb; v[5] = r[1][:100]
    ts = sum(v[1] for v in lines.values()) or 1
    ti = sum(v[0] for v in lines.values()) or 1
    print("\nsource lines of %s with the most warp-stall samples (%d samples, %d warp instructions in all)" % (pat, ts, ti))
    print("  %-22s %6s %7s %7s %7s %7s  %s" % ("file:line", "inst%", "smp%", "barrier", "long_sb", "short_sb", "source"))
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:25]:
        print("  %-22s %5.1f%% %6.1f%% %7d %7d %7d  %s" % ("%s:%d" % (k[0][:16], k[1]), 100 * v[0] / ti, 100 * v[1] / ts, v[2], v[3], v[4], v[5]))


if __name__ == "__main__":
    main()
