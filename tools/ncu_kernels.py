"""Per-kernel table from an ncu csv with time/dram/inst metrics."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
d = collections.defaultdict(lambda: collections.defaultdict(list))
for r in csv.DictReader(lines):
    d[r["Kernel Name"].split("(")[0]][r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
nframes = float(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = 0
for k, m in sorted(d.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = m["gpu__time_duration.sum"]; tot += sum(t)
    g = lambda n: sum(m[n]) / len(t) if m[n] else 0
    print(f"{k:24s} n={len(t):2d} t={sum(t)/len(t)/1e3:7.1f}us rd={g('dram__bytes_read.sum')/1e6:7.2f}MB wr={g('dram__bytes_write.sum')/1e6:6.2f}MB inst={g('smsp__inst_executed.sum')/1e6:6.2f}M warps={g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}%")
print("total per frame us", tot / 1e3 / nframes)
