#!/usr/bin/env python
"""Summarises an .ncu-rep: headline metrics + the instructions with the most stall samples.
Usage: python tools/ncu_top.py report.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for i, h in enumerate(hdr):
    if h in want or "wavefronts_mem_shared" in h or h in ("sm__cycles_elapsed.avg", "sm__cycles_active.avg", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.15):
        print("%-88s %-12s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci = {k: i for i, k in enumerate(h)}
out, tot = [], 0
for r in rows[2:]:
    try:
        v = int(r[ci["Warp Stall Sampling (All Samples)"]])
    except Exception:
        continue
    tot += v
    out.append((v, r[ci["Address"]][-4:], r[ci["Source"]].strip()[:64], r[ci["Instructions Executed"]], r[ci["stall_long_sb"]], r[ci["stall_wait"]],
                r[ci["stall_short_sb"]], r[ci["stall_math"]], r[ci["stall_not_selected"]]))
print("total samples", tot)
print("samples addr instr | executed long wait short math notsel")
for t in sorted(out, reverse=True)[:topn]:
    print(*t)
