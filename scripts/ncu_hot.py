"""Top SASS lines by stall samples from `ncu --page source --csv` (usage: ncu_hot.py file.csv [N])"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), r[isrc].strip(), int(r[iex]), {hdr[i]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for n, (s, src, ex, st) in enumerate(sorted(data, key=lambda d: -d[0])[:N]):
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/tot:5.1f}%  {s:7d}  ex={ex:9d}  {src[:70]:70s} {top}")
