"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0][:60]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("%-62s %5s %12s %11s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %5d %12.1f %11.1f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
