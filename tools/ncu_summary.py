"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of total, launches, avg us."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
    except Exception:
        pass
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
print(f"# total {tot/1e6:.2f} ms over {len(rows)} launches")
print(f"{'share':>7} {'total_ms':>10} {'launches':>8} {'avg_us':>9}  kernel")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{v/tot*100:6.2f}% {v/1e6:10.3f} {c:8d} {v/c/1e3:9.1f}  {n[:100]}")
