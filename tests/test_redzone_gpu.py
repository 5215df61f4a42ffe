"""Out-of-bounds writes: every work buffer of the drivers is allocated between two 256-byte guard bands
(SAT_REDZONE=1, sat_b200/_redzone.py) and the bands must be intact after forward + backward / decode.  Covers the
shapes that select each kernel variant (per-row and grouped attention, scalar / box / row-streamed tensor-core context,
segmented backward, deferred dP, threshold top-k)."""
import pytest
import torch

from oracle import sat_oracle as O
from test_train_forward_gpu import synth

pytestmark = pytest.mark.gpu

TRAIN = [
    (dict(Bi=6, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=9, ragged=True), torch.bfloat16),
    (dict(Bi=5, ncap=1, hw=(5, 3), D=72, A=40, E=24, H=56, V=136, T=7, ragged=True), torch.float32),
    (dict(Bi=3, ncap=5, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=True), torch.bfloat16),
    (dict(Bi=3, ncap=5, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=True), torch.float32),
    (dict(Bi=2, ncap=3, hw=(14, 14), D=1024, A=128, E=256, H=512, V=1000, T=5, ragged=True), torch.bfloat16),
    (dict(Bi=2, ncap=5, hw=(16, 16), D=2048, A=128, E=256, H=1024, V=1000, T=4, ragged=True), torch.bfloat16),
    (dict(Bi=4, ncap=1, hw=(14, 14), D=2048, A=128, E=256, H=1024, V=6400, T=5, ragged=True), torch.bfloat16),
    (dict(Bi=3, ncap=1, hw=(1, 1), D=64, A=32, E=32, H=64, V=128, T=4, ragged=True), torch.bfloat16),
]


@pytest.mark.parametrize("cfg,dtype", TRAIN)
def test_train_step_keeps_guard_bands(cfg, dtype, monkeypatch):
    from sat_b200 import _redzone, decoder
    from sat_b200.packing import PackedWeights
    monkeypatch.setenv("SAT_REDZONE", "1")
    _redzone.violations()                      # drop registrations of earlier tests
    W, ann, caps, lens = synth(**cfg, seed=12)
    fp32 = dtype == torch.float32
    pw = PackedWeights(W, dtype=dtype, device="cuda")
    bld = decoder.annotations_as_bld(ann.cuda(), dtype)
    buf = decoder.train_forward(pw, bld, caps.cuda(), lens.cuda(), 0.1, 1.0, exact=fp32, use_tc=not fp32, backward=True)
    decoder.train_backward(pw, buf)
    torch.cuda.synchronize()
    assert _redzone.violations() == []


DECODE = [
    (dict(B=5, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400), 1, torch.bfloat16),
    (dict(B=5, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400), 5, torch.bfloat16),
    (dict(B=3, hw=(16, 16), D=2048, A=128, E=256, H=512, V=10000), 5, torch.bfloat16),
    (dict(B=3, hw=(14, 14), D=1024, A=128, E=256, H=512, V=1000), 3, torch.bfloat16),
    (dict(B=4, hw=(4, 3), D=64, A=32, E=32, H=64, V=128), 8, torch.float32),
]


@pytest.mark.parametrize("c,k,dtype", DECODE)
def test_decode_keeps_guard_bands(c, k, dtype, monkeypatch):
    from sat_b200 import _redzone, decode, decoder
    monkeypatch.setenv("SAT_REDZONE", "1")
    _redzone.violations()
    W = O.random_weights(c["D"], c["A"], c["E"], c["H"], c["V"], seed=13, sharpen=True)
    g = torch.Generator().manual_seed(14)
    ann = torch.randn(c["B"], c["D"], *c["hw"], generator=g)
    fp32 = dtype == torch.float32
    dw = decode.DecodeWeights(W, dtype, torch.device("cuda"), fp32, not fp32)
    vocab = dict(PAD=0, UNK=c["V"] - 3, START=c["V"] - 2, END=c["V"] - 1)
    t = decode.decode_annotations(dw, decoder.annotations_as_bld(ann.cuda(), dtype), k, 12, 1.0, "LN", 0.5, vocab)
    decode.assemble(t, c["hw"])
    assert _redzone.violations() == []


def test_guard_bands_detect_an_overrun(monkeypatch):
    from sat_b200 import _redzone
    monkeypatch.setenv("SAT_REDZONE", "1")
    _redzone.violations()
    x = _redzone.empty((8,), torch.float32, "cuda", "probe")
    y = _redzone.empty((8,), torch.float32, "cuda", "clean")
    y.fill_(1.0)
    torch.as_strided(x, (9,), (1,))[8] = 1.0          # one element past the end
    assert _redzone.violations() == ["probe"]
