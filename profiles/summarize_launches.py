"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean duration, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"unit us; total {tot:.1f}; launches {sum(v[0] for v in agg.values())}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:4d} avg={t / n:9.2f} share={t / tot:6.3f}")
