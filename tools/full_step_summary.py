"""Summary of an ncu launch list of tools/full_step.py: per-kernel totals of ONE training step (pack_kernel to pack_kernel).
    python tools/full_step_summary.py launches.csv [top-n]"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in csv.DictReader(lines)]
st = [i for i, (n, _) in enumerate(rows) if "pack_kernel" in n]
seg = rows[st[-2]:st[-1]]
agg = collections.defaultdict(lambda: [0, 0.0])
ours = 0.0
OURS = ("tc::", "attention_step", "lstm_bwd", "param_grads", "embed_g", "dP_deferred", "dann_alpha", "colsum", "mean_L", "loss_finalize", "ce_finalize",
        "alpha_sum", "init_state", "tok_init", "ntok", "pack_kernel", "resize_", "cast_captions", "ncap_sum", "dropout_bwd", "ce_rows")
for n, ns in seg:
    k = re.sub(r"\(.*", "", n)[:110]
    agg[k][0] += 1
    agg[k][1] += ns
    if any(o in n for o in OURS):
        ours += ns
tot = sum(v[1] for v in agg.values())
print("one full train step (pack_kernel to pack_kernel): %d launches, %.1f us serialised (cold-cache) device time" % (len(seg), tot / 1e3))
print("libsat_b200 kernels: %.1f us = %.1f%% of the step; the rest is the torchvision trunk on cuDNN / ATen + the optimizer" % (ours / 1e3, 100 * ours / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print("%9.1f us %5.1f%% n=%4d avg %7.1f us  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], v[1] / v[0] / 1e3, k))
