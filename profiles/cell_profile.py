#!/usr/bin/env python
"""The cell kernel's ncu capture as a small JSON keyed by the kernel's source hash.

bench.py reads profiles/cell_kernel_ncu.json for the numbers it cannot measure itself without a
profiler (fp64-pipe utilisation, fp64 instructions and useful lane operations per cell-step, DRAM
bytes per cell-step) and refuses to use them when the kernel sources have changed since the capture
(`kernel_hash`).  Usage, after a gpurun capture
    MHM_CUDA_LAUNCH_LOG=gpurun_out/launches.txt ncu --set full ... -k regex:cell_block_kernel_fast -s N -c 1 \
        -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run here:
    python profiles/cell_profile.py gpurun_out/prof.ncu-rep gpurun_out/launches.txt N > profiles/cell_kernel_ncu.json
"""
import collections
import csv
import hashlib
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["cell_kernel.cuh", "fastmath.cuh", "fastmath_tables.h", "device_types.h",
                  "cell_kernel_launch.inc", "cell_kernel_fast.cu", "Makefile"]


def kernel_hash():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "mhm_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main(rep, launch_log, skip):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    unit = dict(zip(rows[0], rows[1]))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
             "Tbyte": 1e12}

    def f(k):  # times in ms, sizes in bytes, everything else as printed
        return float(d[k].replace(",", "")) * scale.get(unit.get(k, ""), 1.0)

    launches = [l.split() for l in open(launch_log) if l.startswith("cell ")]
    tt_first, n_steps, n_cells, n_members, uniform = (int(x) for x in launches[int(skip)][1:])
    units = float(n_steps) * n_cells * n_members
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h = srows[1]
    ci, ce = h.index("Source"), h.index("Instructions Executed")
    cp = h.index("Predicated-On Thread Instructions Executed")
    warp, lanes = collections.Counter(), collections.Counter()
    for r in srows[2:]:
        try:
            e, t = float(r[ce]), float(r[cp])
        except ValueError:
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ci].strip())
        op = m.group(2) if m else "?"
        warp[op] += e
        lanes[op] += t
    fp64_warp = sum(warp[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
    useful = sum(lanes[k] for k in ("DADD", "DMUL", "DFMA"))
    out = {
        "kernel_hash": kernel_hash(),
        "kernel": d.get("Kernel Name"),
        "device": d.get("device__attribute_display_name"),
        "launch": {"tt_first": tt_first, "steps": n_steps, "cells": n_cells, "members": n_members,
                   "uniform_calendar": uniform, "units": units},
        "duration_ms_under_ncu": f("gpu__time_duration.sum"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lsu_data_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_instructions_per_warp_step": sum(warp.values()) * 32.0 / units,
        "fp64_warp_instructions_per_warp_step": fp64_warp * 32.0 / units,
        "fp64_lane_ops_per_unit": useful / units,
        "dram_bytes_per_unit": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) / units,
        "opcode_share_pct": {k: round(100.0 * v / sum(warp.values()), 2) for k, v in warp.most_common(12)},
    }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--hash":
        print(kernel_hash())
    else:
        main(*sys.argv[1:4])
