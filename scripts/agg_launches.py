"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    n += 1
    if n <= skip:
        continue
    name = row["Kernel Name"].replace("<unnamed>::", "")[:64]
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
print(f"{n - skip} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{100 * t / tot:5.1f}%  {t:10.1f} us  {c:5d} x {t / c:9.2f} us  {k}")
