"""Per-region instruction and stall-sample shares of a kernel from an .ncu-rep source page -- development tool.
usage: ncu_regions.py rep [evals_per_warp_unit]   (regions = maximal runs of SASS instructions with equal execution count)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
W = float(sys.argv[2]) if len(sys.argv) > 2 else 3.125e6
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)      # (a kernel-name line comes first)
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ia = h.index("Instructions Executed"); isrc = h.index("Source"); iss = h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[iss]) for r in data)
segs = []
prev = None
for i, r in enumerate(data):
    c = int(r[ia])
    if prev is None or c != prev:
        segs.append([i, i, c, collections.Counter(), 0])
    segs[-1][1] = i
    t = r[isrc].split(); op = t[1] if t[0].startswith("@") else t[0]
    segs[-1][3][op.split(".")[0]] += 1
    segs[-1][4] += int(r[iss])
    prev = c
print(f"total warp-instrs/W = {tot / W:.1f}, samples = {tots}")
for s in segs:
    n = s[1] - s[0] + 1
    share = n * s[2] / tot * 100
    if share > 0.5 or s[4] / tots > 0.005:
        print(f"{s[0]:5d}-{s[1]:5d} n={n:4d} count/W={s[2] / W:7.2f} instr={share:5.1f}% time={s[4] / tots * 100:5.1f}%  {dict(s[3].most_common(6))}")
