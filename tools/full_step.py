"""BASELINE configs[1] (or --name c3) full training step exactly as bench.py's device-timed loop runs it (encoder + fused decoder
+ backward + Adam), N iterations, for ncu launch lists of the WHOLE step.     python tools/full_step.py [--iters 4]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--name", default="train")
ap.add_argument("--iters", type=int, default=4)
args = ap.parse_args()
from sat_b200.model import SAT  # noqa: E402

c = bench.CFG[args.name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SAT(**bench.hparams(c)).to(dev)
model.encoder.to(memory_format=torch.channels_last)
model.train()
opt = model.configure_optimizers()
img, caps, lens = bench.synth_batch(c["B"], c["T"], c["V"], seed=100, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(args.iters):
    e0.record()
    opt.zero_grad(set_to_none=True)
    loss, aux = model.fused_loss((img.clone(), caps, lens))
    loss.backward()
    opt.step()
    e1.record()
    torch.cuda.synchronize()
    print("iter %d: %.3f ms  loss %.5f" % (it, e0.elapsed_time(e1), float(loss.detach())))
