"""Measures how far bf16 tensor-core decoding (product path) drifts from the fp32 oracle at BASELINE decode dims (C4 greedy,
C5 beam 5): first divergent position per caption and, for greedy, the oracle's log-prob gap between its own choice and the
bf16 path's choice at every step (teacher-forced on the bf16 tokens).  Prints the statistics the floors in
tests/test_full_size_gpu.py are derived from.      python tools/bf16_decode_agreement.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from oracle import sat_oracle as O  # noqa: E402
from test_full_size_gpu import DECODE, _decode, oracle_gaps  # noqa: E402

for name, n in (("C4", 8), ("C5", 4)):
    c = DECODE[name]
    W = O.random_weights(c["D"], c["A"], c["E"], c["H"], c["V"], seed=5, sharpen=True)
    g = torch.Generator().manual_seed(6)
    ann = torch.randn(n, c["D"], *c["hw"], generator=g)
    vocab = dict(PAD=0, UNK=c["V"] - 3, START=c["V"] - 2, END=c["V"] - 1)
    got = _decode(W, ann, c["k"], c["V"])
    f32 = _decode(W, ann, c["k"], c["V"], dtype=torch.float32)
    ref = O.caption(W, ann, vocab, beamk=c["k"], max_gen_length=30, rescore_method="LN")
    print(name, "fp32 CUDA == oracle:", f32[0] == ref[0])
    for i in range(n):
        a, b = got[0][i], ref[0][i]
        m = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
        print("  image %d: len bf16 %d oracle %d, first divergence at %d, score bf16 %.4f oracle %.4f" % (i, len(a), len(b), m, got[1][i], ref[1][i]))
    if c["k"] == 1:
        gaps = oracle_gaps(W, ann, got[0], vocab)
        for i, gp in enumerate(gaps):
            print("  image %d: max oracle gap %.4f nats at step %d; logit range %.1f" % (i, max(gp), gp.index(max(gp)), 0.0))
