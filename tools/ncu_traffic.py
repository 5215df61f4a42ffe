"""Per-kernel DRAM traffic and key counters from an `ncu --set full` report, for bench.py's roofline.traffic:
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ;  python tools/ncu_traffic.py raw.csv profiles/r02_kernel_traffic.json [summary.txt [workload]]
With a workload name (train / c3 / greedy / beam) the result is merged into the JSON under that key and the summary is appended.
Averages over the captured launches of each kernel: dram bytes read + written, duration, warps active, issue active,
tensor-pipe activity, L2 hit rate, registers.  The JSON maps the kernel's base name to {"dram_bytes_per_launch": ...}."""
import collections
import csv
import json
import re
import sys

path, out_json = sys.argv[1], sys.argv[2]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
header = next(rd)
units = next(rd)
col = {n: i for i, n in enumerate(header)}
WANT = {
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__grid_size": "grid", "launch__block_size": "block",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "us": 1.0, "ns": 1e-3, "ms": 1e3}
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for row in rd:
    if len(row) < len(header):
        continue
    name = re.sub(r"\(.*", "", row[col["Kernel Name"]]).strip()
    name = re.sub(r"^void ", "", name)
    for m, key in WANT.items():
        if m in col and row[col[m]] not in ("", "n/a"):
            try:
                v = float(row[col[m]].replace(",", ""))
            except ValueError:
                continue
            v *= SCALE.get(units[col[m]], 1.0)
            agg[name][key].append(v)
res, txt = {}, []
for name, d in agg.items():
    mean = {k: sum(v) / len(v) for k, v in d.items()}
    e = {"launches": len(d.get("duration", [])), "dram_bytes_per_launch": mean.get("dram_read", 0.0) + mean.get("dram_write", 0.0),
         "duration_us": mean.get("duration"), **{k: mean.get(k) for k in ("warps_active_pct", "issue_active_pct", "tensor_pipe_pct", "l2_hit_pct", "dram_pct", "sm_pct", "regs", "grid", "block")}}
    res[name] = e                                   # full name incl. template arguments (bench.py matches by substring)
    txt.append("%-70s n=%d  %.1f us  dram %.2f MB/launch (%.0f%% of peak)  warps active %.0f%%  issue active %.0f%%  tensor pipe %.0f%%  L2 hit %.0f%%  regs %s  grid %s x %s"
               % (name[:70], e["launches"], e["duration_us"] or 0, e["dram_bytes_per_launch"] / 1e6, e["dram_pct"] or 0, e["warps_active_pct"] or 0,
                  e["issue_active_pct"] or 0, e["tensor_pipe_pct"] or 0, e["l2_hit_pct"] or 0, e["regs"], e["grid"], e["block"]))
workload = sys.argv[4] if len(sys.argv) > 4 else None
if workload:
    try:
        full = json.load(open(out_json))
    except Exception:
        full = {}
    full = {k: v for k, v in full.items() if k in ("train", "c3", "greedy", "beam")}     # drop entries of the un-keyed format
    full[workload] = res
    res = full
json.dump(res, open(out_json, "w"), indent=1)
print("\n".join(txt))
if len(sys.argv) > 3:
    with open(sys.argv[3], "a" if workload else "w") as f:
        f.write(("[%s]\n" % workload if workload else "") + "\n".join(txt) + "\n")
