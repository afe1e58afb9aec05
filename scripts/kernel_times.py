"""Per-kernel CUDA time of one steady-state step (torch.profiler / CUPTI, no replay): cheap alternative to the ncu
launch list for day-to-day work.  usage: python scripts/kernel_times.py [out.md]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

dev = torch.device("cuda:0")
BATCH = int(os.environ.get("OCC_BATCH", "1"))  # OCC_BATCH=4 = the bench default
pipe = bench.Pipeline(dev, BATCH)
host = bench.host_inputs(BATCH, 0)
res = {k: v.to(dev) for k, v in host.items()}
for _ in range(3):
    pipe.run(res)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.run(res)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t > 0:
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
out = [f"# per-kernel device time of one step at batch {BATCH} (torch.profiler), total {tot / 1e3:.3f} ms", "", "| kernel | launches | total us | share |",
       "|---|---:|---:|---:|"]
for t, c, k in rows[:45]:
    out.append(f"| {k[:90]} | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
# chronological list (name, us) for order-based attribution
seq = []
for e in prof.events():
    if getattr(e, "device_type", None) is not None and str(e.device_type).endswith("CUDA"):
        t = getattr(e, "device_time", None) or getattr(e, "cuda_time", 0.0)
        seq.append((e.time_range.start, e.name.split("(")[0].replace("void ", "").replace("occ::", "")[:48], t))
seq.sort()
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "kernel_seq.txt" if BATCH == 1 else f"kernel_seq_b{BATCH}.txt"), "w") as f:
    for _, n, t in seq:
        f.write(f"{t:9.1f}  {n}\n")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")
