"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/launches_summary.txt"""
import collections
import csv
import io
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "")
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"):
        v *= 1e3
    elif r["Metric Unit"] in ("ms", "msecond"):
        v *= 1e6
    agg[n][0] += 1
    agg[n][1] += v / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms total device time (cold-cache, serialised: compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} x{v[0]:5d} {v[1]:9.2f} ms {100 * v[1] / tot:5.1f}%")
