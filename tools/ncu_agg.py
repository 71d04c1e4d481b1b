"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) per kernel name: count, median, min (us)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][-64:]
    agg.setdefault(name, []).append(float(r[-1].replace(",", "")))
tot = 0.0
for k, v in agg.items():
    v2 = sorted(v)
    med = v2[len(v2) // 2] / 1e3
    tot += med
    print(f"{k:64s} n={len(v):3d} median={med:9.2f} us  min={v2[0] / 1e3:9.2f}")
print(f"sum of medians: {tot:.1f} us")
