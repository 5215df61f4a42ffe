"""Soak: many alternating training steps and captioning calls on one module (bf16, real resnet18 trunk); checks that device memory
does not grow and results stay finite.     python tools/soak.py [--iters 300]"""
import argparse
import os
import sys
import warnings

warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from sat_b200.model import SAT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=300)
args = ap.parse_args()
c = dict(bench.CFG["train"], arch="resnet18", B=32)
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SAT(**bench.hparams(c, dropout=0.2, embedding_dropout=0.1, label_smoothing=0.1, decoder_layers=2)).to(dev)
m.encoder.to(memory_format=torch.channels_last)
opt = m.configure_optimizers()
img, caps, lens = bench.synth_batch(c["B"], c["T"], c["V"], seed=1, device=dev)
mem = []
for it in range(args.iters):
    m.train()
    out = m.training_step((img.clone(), caps, lens), it)
    opt.zero_grad(set_to_none=True)
    out["loss"].backward()
    opt.step()
    if it % 10 == 0:
        r = m.caption(img[:8].clone(), beamk=3 if it % 20 else 1, max_gen_length=12,
                      sample_method="beam" if it % 30 else "multinomial")
        assert len(r[0]) == 8
    if it % 50 == 0 or it == args.iters - 1:
        torch.cuda.synchronize()
        mem.append(torch.cuda.memory_allocated())
        print("iter %d loss %.4f allocated %.1f MB bad_tokens %.0f" % (it, float(out["loss"]), mem[-1] / 1e6, float(out["bad_token_ids"])))
        assert torch.isfinite(out["loss"])
assert mem[-1] <= mem[1] * 1.02 + (1 << 20), "device memory grew: %s" % mem
print("soak ok: loss %.4f -> %.4f" % (9.0, float(out["loss"])))
