"""GPU parity of the fused teacher-forced forward + loss against the CPU oracle and the committed
reference goldens (fp32: 1e-5 relative on alpha / logits / loss; bf16: 2e-2)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import sat_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_cuda_forward(W, ann, caps, lens, ls, gamma, dtype=torch.float32, exact=True, use_tc=False, logits_f32=True):
    from sat_b200.packing import PackedWeights
    from sat_b200 import decoder
    pw = PackedWeights(W, dtype=dtype, device="cuda")
    bld = decoder.annotations_as_bld(ann.cuda(), dtype)
    buf = decoder.train_forward(pw, bld, caps.cuda(), lens.cuda(), ls, gamma, exact=exact, use_tc=use_tc,
                                logits_f32=logits_f32, backward=False)
    torch.cuda.synchronize()
    out = buf.t["out"].cpu()
    logits = buf.t["logits"].float().permute(1, 0, 2).cpu()        # [B,T,V]
    return dict(loss=float(out[0]), ce=float(out[1]), reg=float(out[2]), acc=float(out[3]), logits=logits,
                alphas=buf.t["alphas"].cpu(), buf=buf)


@pytest.mark.parametrize("name", ["train_small", "train_ragged"])
def test_forward_fp32_vs_reference_golden(name):
    z, W, _ = load_golden(name)
    ann = torch.from_numpy(z["ann"])
    caps, lens = torch.from_numpy(z["caps"]), torch.from_numpy(z["lengths"])
    r = run_cuda_forward(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    assert relerr(r["alphas"], z["alphas"]) < 1e-5
    assert relerr(r["logits"], z["logits"]) < 1e-5
    assert abs(r["loss"] - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert abs(r["acc"] - float(z["acc"])) < 1e-6


def synth(Bi, ncap, hw, D, A, E, H, V, T, ragged, seed=0, sharpen=False, layers=1):
    W = O.random_weights(D, A, E, H, V, seed=seed, sharpen=sharpen, layers=layers)
    g = torch.Generator().manual_seed(seed + 100)
    ann = torch.randn(Bi, D, hw[0], hw[1], generator=g)
    caps = torch.randint(1, V - 3, (Bi, ncap, T + 1), generator=g)
    caps[:, :, 0] = V - 2
    lens = torch.randint(2, T + 1, (Bi, ncap), generator=g) if ragged else torch.full((Bi, ncap), T)
    return W, ann, caps, lens


@pytest.mark.parametrize("cfg", [
    dict(Bi=8, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=20, ragged=False),   # BASELINE config 1 decoder shape
    dict(Bi=6, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=12, ragged=True),
    dict(Bi=3, ncap=5, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=9, ragged=True),        # COCO-like 5 captions / image
    dict(Bi=4, ncap=1, hw=(16, 16), D=2048, A=128, E=256, H=512, V=10000, T=6, ragged=True),    # config 5 dims
    dict(Bi=5, ncap=1, hw=(5, 3), D=72, A=40, E=24, H=56, V=136, T=7, ragged=True),             # odd (multiple-of-8) dims
])
def test_forward_fp32_vs_oracle(cfg):
    W, ann, caps, lens = synth(**cfg)
    ref = O.train_loss(W, ann, caps, lens, label_smoothing=0.1, att_gamma=1.0)
    r = run_cuda_forward(W, ann, caps, lens, 0.1, 1.0)
    assert relerr(r["alphas"], ref["alphas"]) < 1e-5
    assert relerr(r["logits"], ref["logits"]) < 1e-5
    assert abs(r["loss"] - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    assert abs(r["acc"] - float(ref["acc"])) < 1e-6


@pytest.mark.parametrize("use_tc", [False, True])
def test_forward_bf16_vs_oracle(use_tc):
    cfg = dict(Bi=8, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=20, ragged=True)
    W, ann, caps, lens = synth(**cfg)
    ref = O.train_loss(W, ann, caps, lens, label_smoothing=0.0, att_gamma=1.0)
    r = run_cuda_forward(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=use_tc, logits_f32=False)
    assert relerr(r["logits"], ref["logits"]) < 2e-2
    assert abs(r["loss"] - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))
    assert relerr(r["alphas"], ref["alphas"]) < 2e-2


@pytest.mark.parametrize("cfg", [
    dict(Bi=3, ncap=5, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=9, ragged=True),
    dict(Bi=2, ncap=5, hw=(16, 16), D=2048, A=128, E=256, H=512, V=1000, T=5, ragged=True),     # config 5 tile, L=256
    dict(Bi=2, ncap=8, hw=(14, 14), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=True),      # L=196: boxes overrun the image
    dict(Bi=3, ncap=2, hw=(14, 14), D=512, A=96, E=256, H=512, V=1000, T=6, ragged=False),
    dict(Bi=2, ncap=3, hw=(14, 14), D=1024, A=128, E=256, H=512, V=1000, T=5, ragged=True),      # row-streamed kernel, partial last stage
    dict(Bi=2, ncap=3, hw=(18, 18), D=1024, A=128, E=256, H=512, V=1000, T=4, ragged=True),      # L = 324 > 256: row-streamed only
    dict(Bi=2, ncap=3, hw=(18, 18), D=512, A=128, E=256, H=512, V=1000, T=4, ragged=True),       # L = 324, D = 512: scalar grouped kernel
])
def test_forward_bf16_multi_caption_vs_oracle(cfg):
    """several caption rows per image in bf16: the grouped attention kernel (tensor-core context, alpha rounded to bf16)."""
    W, ann, caps, lens = synth(**cfg)
    ref = O.train_loss(W, ann, caps, lens, label_smoothing=0.0, att_gamma=1.0)
    r = run_cuda_forward(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=True, logits_f32=False)
    assert relerr(r["logits"], ref["logits"]) < 2e-2
    assert abs(r["loss"] - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))
    assert relerr(r["alphas"], ref["alphas"]) < 2e-2


@pytest.mark.parametrize("cfg,lens_override", [
    (dict(Bi=1, ncap=1, hw=(2, 2), D=64, A=32, E=32, H=64, V=128, T=5, ragged=False), None),          # single caption, L=4
    (dict(Bi=3, ncap=1, hw=(1, 1), D=64, A=32, E=32, H=64, V=128, T=4, ragged=False), [[1], [4], [2]]),   # L=1, shortest length 1
    (dict(Bi=2, ncap=3, hw=(3, 5), D=128, A=64, E=64, H=128, V=256, T=6, ragged=False), [[6, 1, 3], [2, 6, 6]]),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_edge_shapes(cfg, lens_override, dtype):
    """ragged / minimal inputs: one caption, one location, length-1 captions, several captions per image."""
    W, ann, caps, lens = synth(**cfg)
    if lens_override is not None:
        lens = torch.tensor(lens_override)
    ref = O.train_loss(W, ann, caps, lens, label_smoothing=0.1, att_gamma=1.0)
    fp32 = dtype == torch.float32
    r = run_cuda_forward(W, ann, caps, lens, 0.1, 1.0, dtype=dtype, exact=fp32, use_tc=not fp32, logits_f32=fp32)
    tol = 1e-5 if fp32 else 2e-2
    assert relerr(r["alphas"], ref["alphas"]) < tol
    assert relerr(r["logits"], ref["logits"]) < tol
    assert abs(r["loss"] - float(ref["loss"])) < tol * abs(float(ref["loss"]))
    # rows past their length are exactly zero, like the reference's pre-zeroed buffers (model.py:504-506)
    B, T = r["logits"].shape[:2]
    fl = lens.reshape(-1)
    for b in range(B):
        assert float(r["logits"][b, int(fl[b]):].abs().max() if int(fl[b]) < T else 0.0) == 0.0
        assert float(r["alphas"][b, int(fl[b]):].abs().max() if int(fl[b]) < T else 0.0) == 0.0
