"""GPU parity of the hand-written BPTT against autograd of the reference (golden) / the CPU oracle."""
import pytest
import torch

from conftest import load_golden
from oracle import sat_oracle as O
from test_train_forward_gpu import relerr, synth

pytestmark = pytest.mark.gpu


def run_cuda_fwd_bwd(W, ann, caps, lens, ls, gamma, dtype=torch.float32, exact=True, use_tc=False):
    from sat_b200.packing import PackedWeights
    from sat_b200 import decoder
    pw = PackedWeights(W, dtype=dtype, device="cuda")
    bld = decoder.annotations_as_bld(ann.cuda(), dtype)
    buf = decoder.train_forward(pw, bld, caps.cuda(), lens.cuda(), ls, gamma, exact=exact, use_tc=use_tc,
                                logits_f32=False, backward=True)
    G, d_ann = decoder.train_backward(pw, buf)
    torch.cuda.synchronize()
    Bi, D, h, w = ann.shape
    d_ann = d_ann.float().reshape(Bi, h, w, D).permute(0, 3, 1, 2).cpu()
    return float(buf.t["out"][0]), {k: v.cpu() for k, v in G.items()}, d_ann


@pytest.mark.parametrize("name", ["train_small", "train_ragged"])
def test_backward_fp32_vs_reference_golden(name):
    z, W, Gref = load_golden(name)
    ann = torch.from_numpy(z["ann"])
    caps, lens = torch.from_numpy(z["caps"]), torch.from_numpy(z["lengths"])
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    assert abs(loss - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    for k, g in Gref.items():
        assert relerr(G[k], g) < 2e-5, k
    assert relerr(d_ann, z["d_ann"]) < 2e-5


def oracle_grads(W, ann, caps, lens, ls, gamma):
    Wg = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    a = ann.clone().requires_grad_(True)
    r = O.train_loss(Wg, a, caps, lens, ls, gamma)
    r["loss"].backward()
    return float(r["loss"]), {k: v.grad for k, v in Wg.items()}, a.grad


@pytest.mark.parametrize("cfg", [
    dict(Bi=6, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=8, ragged=True),
    dict(Bi=3, ncap=5, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=True),
    dict(Bi=5, ncap=1, hw=(5, 3), D=72, A=40, E=24, H=56, V=136, T=7, ragged=True),
])
def test_backward_fp32_vs_oracle(cfg):
    W, ann, caps, lens = synth(**cfg)
    W["attention.f_att.weight"] *= 10      # non-uniform alpha so the softmax backward is exercised
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.1, 1.0)
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, 0.1, 1.0)
    assert abs(loss - loss_ref) < 1e-5 * abs(loss_ref)
    for k, g in Gref.items():
        assert relerr(G[k], g) < 5e-5, k
    assert relerr(d_ann, da_ref) < 5e-5


@pytest.mark.parametrize("use_tc", [False, True])
def test_backward_bf16_vs_oracle(use_tc):
    cfg = dict(Bi=8, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=10, ragged=True)
    W, ann, caps, lens = synth(**cfg)
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.0, 1.0)
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=use_tc)
    assert abs(loss - loss_ref) < 2e-2 * abs(loss_ref)
    for k, g in Gref.items():
        assert relerr(G[k], g) < 6e-2, k
    assert relerr(d_ann, da_ref) < 6e-2
