#!/usr/bin/env python
"""Turn an .ncu-rep (captured on the B200 box under gpurun) into the text summary committed
under profiles/.  Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), " device:", d.get("device__attribute_display_name"))
        for k in KEYS:
            if k in d:
                print("  %-82s %s %s" % (k, d[k], u.get(k, "")))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        h = rows[1]
        ci, ce, ct, cs = (h.index("Source"), h.index("Instructions Executed"),
                          h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"))
        import collections
        import re
        ops, thr, st = collections.Counter(), collections.Counter(), collections.Counter()
        for r in rows[2:]:
            try:
                e, t, s = float(r[ce]), float(r[ct]), float(r[cs])
            except ValueError:
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci].strip())
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += e
            thr[op] += t
            st[op] += s
        tot, ts = sum(ops.values()), max(1.0, sum(st.values()))
        print("  SASS opcode mix (share of warp instructions, mean active lanes, share of stall samples):")
        for op, e in ops.most_common(14):
            print("    %-8s %5.1f%%  lanes %4.1f  stalls %5.1f%%" % (op, 100 * e / tot, thr[op] / e, 100 * st[op] / ts))


if __name__ == "__main__":
    main(sys.argv[1])
