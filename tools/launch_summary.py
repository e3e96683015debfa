#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv [launches_per_plan]"""
import collections, csv, re, sys
path = sys.argv[1]; per = int(sys.argv[2]) if len(sys.argv) > 2 else 67
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))[-per:]
agg, tot = collections.OrderedDict(), 0.0
def us(row):
    t = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    return t / 1000 if u == "ns" else t * 1000 if u == "ms" else t
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void m3pc::<unnamed>::|^m3pc::<unnamed>::", "", name)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us(row); tot += us(row)
print(f"# {path}: last {per} launches (one plan), serialised cold-cache ncu times; total {tot:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{n:3d}  {k}")
if "-v" in sys.argv:
    for i, row in enumerate(rows):
        name = re.sub(r"^void m3pc::<unnamed>::|^m3pc::<unnamed>::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        print(f"{i:3d} {name[:44]:44s} {us(row):8.1f} us grid {row['Grid Size']}")
