#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, median / max / total ms, share.
Usage: python tools/ncu_summary.py gpurun_out/launches.csv"""
import csv, collections, re, statistics, sys
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        name = re.sub(r"^kzgb200::", "", r["Kernel Name"]).split("(")[0]
        rows.append((name[:36], float(r["Metric Value"].replace(",", "")) / 1e6))
by = collections.OrderedDict()
for n, ms in rows:
    by.setdefault(n, []).append(ms)
verify = {n: v for n, v in by.items() if not re.match(r"lag_|quotient|blob_scalars|setup_tables|void at::|harness", n)}
tot = sum(sum(v) for v in verify.values())
print("%-36s %9s %10s %10s %10s %8s" % ("kernel", "launches", "median_ms", "max_ms", "total_ms", "share"))
for n, v in sorted(verify.items(), key=lambda kv: -sum(kv[1])):
    print("%-36s %9d %10.4f %10.4f %10.3f %7.2f%%" % (n, len(v), statistics.median(v), max(v), sum(v), 100 * sum(v) / tot))
print("\n# workload generation / setup (outside the timed region)")
for n, v in by.items():
    if n not in verify:
        print("%-36s %9d %10.4f %10.4f %10.3f" % (n, len(v), statistics.median(v), max(v), sum(v)))
