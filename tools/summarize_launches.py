"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total, share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row.get("Kernel Name", ""))
    name = re.sub(r"^void ", "", name)
    short = re.sub(r"<.*", "", name)
    m = re.search(r"genz::(\w+)<\(?int\)?(\d+)>", name)
    if m:
        short += f"<{m.group(1)}<{m.group(2)}>>"
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row["Metric Unit"]
    v *= {"ms": 1e3, "ns": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1.0)
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
      f"(ncu per-launch times: cold cache, serialised -- compare shares, not absolutes)")
print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {v[0]:8d} {v[1] / 1e3:10.2f} {v[1] / v[0]:9.1f} {100 * v[1] / tot:6.2f}%")
