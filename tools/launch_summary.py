"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share over the last N launches."""
import collections
import csv
import sys

path, last = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 120
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        rows.append((r["Kernel Name"].split("(")[0], v if r["Metric Unit"] == "ns" else v * 1e3))
agg = collections.defaultdict(list)
for name, v in rows[-last:]:
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{len(rows)} launches in file; last {last}: total {tot / 1e3:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:22s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:8.1f} us  share={100 * sum(v) / tot:5.1f}%")
