"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one steady-state step of bench.py
(delimited by consecutive vp_index_geom_kernel launches), aggregated per kernel.  usage: summarize_launches.py csv [out.md]"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("occ::", "")
    m = re.match(r"at::native::.*?(\w+_kernel\w*|\w+Functor\w*)", name)
    if name.startswith("at::") or "at::native" in name:
        f = re.findall(r"([A-Za-z_0-9]+(?:Functor|kernel|Kernel)[A-Za-z_0-9]*)", name)
        return "torch:" + (f[-1] if f else name[:40])
    return name[:70]


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append((int(r["ID"]), short(r["Kernel Name"]), float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    marks = [i for i, r in enumerate(rows) if (r[1].startswith("vp_index_geom_kernel") or r[1].startswith("lift_front_kernel"))]
    if len(marks) >= 2:
        lo, hi = marks[-2], marks[-1]
    else:
        lo, hi = 0, len(rows)
    # a step starts a few torch geometry kernels before vp_index; close enough for shares
    step = rows[lo:hi]
    agg = OrderedDict()
    for _, n, t, g, b in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    out = [f"# launch list summary: {path}", "",
           f"one steady-state step = launches {rows[lo][0]}..{rows[hi - 1][0]} ({len(step)} launches), "
           f"sum of kernel durations {total / 1e6:.3f} ms (ncu-serialised, cold cache: compare SHARES)", "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {n} | {c} | {t / 1e3:.1f} | {100 * t / total:.1f}% |")
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
