"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> markdown table."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        val *= 1e3
    elif unit in ("ms", "msecond"):
        val *= 1e6
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = short.replace("(anonymous namespace)::", "")
    agg[short][0] += 1
    agg[short][1] += val
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k[:110], n, t / 1e6, 100 * t / tot))
print("| **total** | %d | %.3f | 100%% |" % (sum(v[0] for v in agg.values()), tot / 1e6))
