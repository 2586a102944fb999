"""Top SASS lines of one kernel from `ncu --page source --csv`: instructions executed, stall samples."""
import csv, sys
path, kernel_sub, topn = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
mode = sys.argv[4] if len(sys.argv) > 4 else "inst"
rows = list(csv.reader(open(path)))
out, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and kernel_sub in cur and hdr and len(r) >= len(hdr) - 2:
        d = dict(zip(hdr, r))
        out.append(d)
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot_i = sum(num(d["Instructions Executed"]) for d in out); tot_s = sum(num(d["# Samples"]) for d in out)
print(f"{len(out)} SASS lines, {tot_i/1e6:.2f}M warp-instructions, {tot_s:.0f} samples")
key = "Instructions Executed" if mode == "inst" else "# Samples"
if mode == "list":
    for i, d in enumerate(out):
        print(f"{i:4d} {num(d['Instructions Executed'])/1e3:9.1f}k {num(d['# Samples']):6.0f} thr={d['Avg. Threads Executed']:>5s} {d['Source'][:110]}")
else:
    for d in sorted(out, key=lambda d: -num(d[key]))[:topn]:
        print(f"{num(d['Instructions Executed'])/1e3:9.1f}k  samples={num(d['# Samples']):6.0f}  thr={d['Avg. Threads Executed']:>5s}  long_sb={d.get('stall_long_sb','')}  short={d.get('stall_short_sb','')} wait={d.get('stall_wait','')}  {d['Source'][:100]}")
