"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST iteration.
    python tools/launch_summary.py launches.csv <name-substring-of-the-first-kernel-of-an-iteration> [top-n] [raw-out.csv]
The optional raw output holds one line per launch of that iteration (launch, kernel, grid, block, duration_ns).
"""
import collections
import csv
import re
import sys

path, marker = sys.argv[1], sys.argv[2]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = list(csv.DictReader(lines))
rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in recs]
starts = [i for i, (nm, _) in enumerate(rows) if marker in nm]
last = rows[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
for nm, ns in last:
    k = re.sub(r"\(.*", "", nm)[:100]
    agg[k][0] += 1
    agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
print("last iteration: %d launches, %.1f us summed kernel time (ncu: serialised, cold caches)" % (len(last), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%8.1f us %5.1f%% n=%4d avg %7.1f us  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], v[1] / v[0] / 1e3, k))
if len(sys.argv) > 4:
    with open(sys.argv[4], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "grid", "block", "duration_ns"])
        for i, x in enumerate(recs[starts[-1]:]):
            w.writerow([i, re.sub(r"\(.*", "", x["Kernel Name"])[:120], x["Grid Size"], x["Block Size"], int(float(x["Metric Value"].replace(",", "")))])
