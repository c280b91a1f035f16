"""Summarise an ncu launch list (gpu__time_duration per kernel) into one training step's kernel budget.
    python benchmarks/launch_summary.py gpurun_out/launches.csv [min_us]
"""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        L.append((r[ki], float(r[vi].replace(",", ""))))
    except ValueError:
        pass
idx = [i for i, (n, t) in enumerate(L) if "softmax_xent_kernel" in n]
a, b = idx[-2], idx[-1]
step = L[a:b]
print(f"launches/step {len(step)}  sum {sum(t for n, t in step) / 1000:.1f} us")
agg = collections.OrderedDict()
for n, t in step:
    n = re.sub(r"\(.*", "", n.replace("void ", ""))[:70]
    agg.setdefault(n, [0, 0])
    agg[n][0] += t / 1000
    agg[n][1] += 1
for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{t:8.1f} us x{c:3d}  {n}")
print()
for n, t in step:
    if t / 1000 >= min_us:
        print(f"{t / 1000:8.1f}", re.sub(r"\(psd::Tmap.*", "", n)[:110])
