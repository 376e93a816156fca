"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [clean_copy.csv]"""
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
if len(sys.argv) > 2:
    open(sys.argv[2], "w").writelines(lines)
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = {}
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r["Kernel Name"][:70]
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + ms)
tot = sum(t for _, t in agg.values())
print(f"# {sum(n for n, _ in agg.values())} launches, {tot:.2f} ms total")
print("kernel | launches | total ms | share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k} | {n} | {t:.3f} | {100 * t / tot:.1f}%")
