"""Probe: BASELINE configs[1] training step (bench.py's `train` workload) eager vs replayed from ONE CUDA graph
(encoder forward + fused decoder + backward + Adam captured whole).     python tools/graph_probe.py [--name c3]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--name", default="train")
ap.add_argument("--steps", type=int, default=8)
args = ap.parse_args()
from sat_b200.model import SAT  # noqa: E402

c = bench.CFG[args.name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SAT(**bench.hparams(c)).to(dev)
model.encoder.to(memory_format=torch.channels_last)
model.train()
opt = model.configure_optimizers()
for g in opt.param_groups:
    g["capturable"] = True
img, caps, lens = bench.synth_batch(c["B"], c["T"], c["V"], seed=100, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img, caps, lens))
    loss.backward()
    opt.step()
    return loss


def timeit(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = 1e3 * (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        l = step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print("eager: %.3f ms/step device, %.3f ms host issue; loss %.5f" % (*timeit(step, args.steps), float(l)))
graph = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(graph):
    static_loss = step()
torch.cuda.synchronize()
graph.replay()
torch.cuda.synchronize()
print("graph: %.3f ms/step device, %.3f ms host issue; loss %.5f" % (*timeit(graph.replay, args.steps), float(static_loss)))
