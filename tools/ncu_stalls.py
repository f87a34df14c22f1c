"""Per-region stall-reason breakdown of a kernel from an .ncu-rep source page -- development tool.
usage: ncu_stalls.py rep [units] [lo hi]   (regions as in ncu_regions.py; with lo hi: per-instruction listing of that range)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5625e6
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)      # (a kernel-name line comes first)
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ia = h.index("Instructions Executed"); isrc = h.index("Source"); iss = h.index("Warp Stall Sampling (All Samples)")
reasons = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tots = sum(int(r[iss]) for r in data)
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, hi + 1):
        r = data[i]
        top = sorted(((int(r[k]), n) for k, n in reasons if int(r[k])), reverse=True)[:3]
        print(i, r[ia], r[iss], r[isrc][:60].ljust(60), " ".join(f"{n}:{v}" for v, n in top))
    sys.exit(0)
segs = []
prev = None
for i, r in enumerate(data):
    c = int(r[ia])
    if prev is None or c != prev:
        segs.append([i, i, c, collections.Counter(), 0])
    segs[-1][1] = i
    for k, n in reasons:
        segs[-1][3][n] += int(r[k])
    segs[-1][4] += int(r[iss])
    prev = c
tot_r = collections.Counter()
for s in segs:
    tot_r.update(s[3])
print("all:", {k: f"{v / tots * 100:.1f}%" for k, v in tot_r.most_common(8)})
for s in segs:
    if s[4] / tots > 0.008:
        n = s[1] - s[0] + 1
        print(f"{s[0]:5d}-{s[1]:5d} n={n:4d} count/W={s[2] / W:7.2f} time={s[4] / tots * 100:5.1f}%  " +
              " ".join(f"{k}:{v / max(s[4], 1) * 100:.0f}%" for k, v in s[3].most_common(5)))
