"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log per kernel.
usage: ncu_times.py launches.csv"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
agg, cnt = defaultdict(lambda: defaultdict(float)), defaultdict(int)
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("void ", "")
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[k][r[mi]] += v
    if r[mi] == "gpu__time_duration.sum":
        cnt[k] += 1
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print(f"{k[:44]:44s} n={cnt[k]:3d} ms={d['gpu__time_duration.sum'] / 1e6:8.3f} "
          f"rdGB={d['dram__bytes_read.sum'] / 1e9:7.2f} wrGB={d['dram__bytes_write.sum'] / 1e9:7.2f}")
