"""Summarise an .ncu-rep (raw + source pages) -- development tool.  usage: ncu_summary.py rep [n_warp_iters]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
for k in want:
    if k in d:
        print(f"{k:90s} {d[k][0]} {d[k][1]}")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(d[h][0])
        if v > 0.05:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]; data = rows[2:]
ia = h2.index("Instructions Executed"); isrc = h2.index("Source")
by = collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    by[op.split(".")[0]] += int(r[ia])
tot = sum(by.values())
print("SASS instrs:", len(data), "executed warp-instrs:", tot)
print("mix:", ", ".join(f"{op} {c / tot * 100:.1f}%" for op, c in by.most_common(18)))
if len(sys.argv) > 2:
    wi = float(sys.argv[2])
    print("per warp-iteration:", ", ".join(f"{op} {c / wi:.1f}" for op, c in by.most_common(18)), " total", tot / wi)
