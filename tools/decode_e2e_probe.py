"""Where does the end-to-end greedy / beam captioning call (SAT.caption_stream, bench.py's e2e leg) spend its time per 256-image chunk?
    python tools/decode_e2e_probe.py [--name greedy|beam]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from sat_b200 import decode, decoder  # noqa: E402
from sat_b200.model import SAT  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--name", default="greedy")
args = ap.parse_args()
c = bench.CFG[args.name]
dev = torch.device("cuda")
torch.manual_seed(0)
model = SAT(**bench.hparams(c)).to(dev).eval()
model.encoder.to(memory_format=torch.channels_last)
img_h = torch.rand(256, 3, 224, 224).pin_memory()
dw = decode.inference_weights(model)
voc = dict(PAD=0, UNK=c["V"] - 3, START=c["V"] - 2, END=c["V"] - 1)


def timeit(fn, n=6, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    host = 1e3 * (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n, host


print("H2D 256 images: %.2f ms (host issue %.2f)" % timeit(lambda: img_h.to(dev, non_blocking=True)))
img_d = img_h.to(dev)
with torch.no_grad():
    print("encode 256 (eval): %.2f ms (host issue %.2f)" % timeit(lambda: model.encode(img_d.clone())))
    ann = model.encode(img_d.clone())
bld = decoder.annotations_as_bld(ann, dw.pw.dtype)
print("decode 256 k=%d: %.2f ms (host issue %.2f)" % ((c["k"],) + timeit(lambda: decode.decode_annotations(dw, bld, c["k"], c["S"], 1.0, None, 0.5, voc))))
t = decode.decode_annotations(dw, bld, c["k"], c["S"], 1.0, None, 0.5, voc)
torch.cuda.synchronize()
print("assemble (D2H + python lists): %.2f ms (host %.2f)" % timeit(lambda: decode.assemble(t, tuple(ann.shape[2:]), return_all=False)))
chunks = lambda: (img_h for _ in range(4))
for rep in range(3):
    print("caption_stream, 4 chunks: %.2f ms per chunk" % (timeit(lambda: list(model.caption_stream(chunks(), beamk=c["k"], max_gen_length=c["S"])), n=3, w=1)[0] / 4))
