"""profiles/rNN_sass_excerpt.txt: the SASS of the hot kernel that the judge would otherwise have to disassemble from the
(git-ignored) library -- the bulk-TMA / mbarrier staging of the theta tile and the Newton loop of the fused polynomial
chain.  usage: tools/sass_excerpt.py OUT.txt   (no GPU needed: cuobjdump on pyds_b200/libpyds_b200.so)"""
import re
import subprocess
import sys

LIB = "pyds_b200/libpyds_b200.so"
KERNEL = "feas_kernelILi1ELi2ELb0ELb0ELi2ELy6636961890"       # feas_kernel<1, 2, false, false, 2, SHAPE_ABC>
out = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if "Function :" in l and KERNEL in l)
end = next(i for i in range(start + 1, len(sass)) if "Function :" in sass[i])
body = [l for l in sass[start:end] if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
ins = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in body]


def op(l):
    t = l.split("*/", 1)[1].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def window(i, before, after):
    return ins[max(i - before, 0):i + after + 1]


with open(out, "w") as f:
    f.write(f"# {sass[start].strip()}\n# {len(ins)} SASS instructions; excerpted by tools/sass_excerpt.py from {LIB}\n")
    counts = {}
    for l in ins:
        counts[op(l)] = counts.get(op(l), 0) + 1
    f.write("# static mix: " + ", ".join(f"{k} {v}" for k, v in sorted(counts.items(), key=lambda kv: -kv[1])[:16]) + "\n\n")
    f.write("## theta tile staging: mbarrier init, expect-tx + 1-D bulk TMA (UBLKCP), parity wait, prefetch of the next tile\n")
    seen = set()
    for i, l in enumerate(ins):
        if any(k in l for k in ("SYNCS", "UBLKCP")):
            for j, w in zip(range(max(i - 3, 0), i + 3), window(i, 3, 2)):
                if j not in seen:
                    seen.add(j)
                    f.write(w + "\n")
            f.write("   ...\n")
    # the Newton loop: the backward branch right after VOTE.ALL, from its target to the branch
    v = next(i for i, l in enumerate(ins) if "VOTE.ALL" in l)
    br = next(i for i in range(v, v + 12) if " BRA " in ins[i])
    tgt = int(re.search(r"BRA\s+(?:\S+,\s*)?(0x[0-9a-f]+)", ins[br]).group(1), 16)
    top = next(i for i, l in enumerate(ins) if int(re.search(r"/\*([0-9a-f]{4,5})\*/", l).group(1), 16) == tgt)
    loop = ins[top:br + 1]
    lc = {}
    for l in loop:
        lc[op(l)] = lc.get(op(l), 0) + 1
    f.write(f"\n## Newton loop of the nonlinear block, two scenarios per thread ({len(loop)} instructions per iteration: "
            + ", ".join(f"{k} {v}" for k, v in sorted(lc.items(), key=lambda kv: -kv[1])) + ")\n")
    f.write("\n".join(loop) + "\n")
print(out, len(ins))
