"""Opcode mix and hottest SASS lines of an .ncu-rep (source page): usage: ncu_hot_sass.py rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
si, ie, ss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
agg, tot, lines = {}, 0, []
for k, r in enumerate(rows[2:]):
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    t = r[si].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
    op = op.split(".")[0]
    agg[op] = agg.get(op, 0) + n
    tot += n
    lines.append((int(r[ss] or 0), n, k, r[si].strip()))
print("total warp instructions", tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:18]:
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%")
smp = sum(l[0] for l in lines)
print("hottest lines by stall samples (samples, executed, index, sass)")
for l in sorted(lines, reverse=True)[:top]:
    print(f"{100 * l[0] / max(smp, 1):5.1f}% {l[1]:10d} {l[2]:5d}  {l[3][:100]}")
