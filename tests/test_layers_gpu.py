"""GPU: decoder_layers > 1 (nn.LSTM num_layers, model.py:175-180).  The states are [layers,B,H]; layer l > 0 takes the new
hidden state of layer l-1; attention, the beta gate and the output layer read the top layer (model.py:299-300,327,538-547);
init_lstm.init is [2*layers*H, E] and its output is reinterpreted as [2*layers, B, H] (model.py:79-80).  Checked against
goldens generated from the unmodified reference (oracle/make_golden.py layers) and against the CPU oracle."""
import warnings

import pytest
import torch
from torch import nn

from conftest import load_golden
from oracle import ref_harness as rh
from oracle import sat_oracle as O
from test_decode_gpu import VOC, cuda_caption
from test_train_backward_gpu import oracle_grads, run_cuda_fwd_bwd
from test_train_forward_gpu import relerr, run_cuda_forward, synth

warnings.filterwarnings("ignore")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["train_layers2", "train_layers3"])
def test_train_vs_reference_golden(name):
    """train_layers2: two layers, D=16 A=8 E=10 H=14 V=50 (zero-padded storage), 2 captions per image, ragged;
    train_layers3: three layers, tile-sized dims, ragged"""
    z, W, Gref = load_golden(name)
    ann = torch.from_numpy(z["ann"])
    caps, lens = torch.from_numpy(z["caps"]), torch.from_numpy(z["lengths"])
    V0 = int(z["dims"][4])
    r = run_cuda_forward(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    assert relerr(r["alphas"], z["alphas"]) < 1e-5
    assert relerr(r["logits"][..., :V0], z["logits"]) < 1e-5
    assert abs(r["loss"] - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert abs(r["acc"] - float(z["acc"])) < 1e-6
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    assert abs(loss - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert any(k.endswith("_l1") for k in Gref)
    for k, g in Gref.items():
        assert tuple(G[k].shape) == tuple(g.shape), k
        assert relerr(G[k], g) < 2e-5, k
    assert relerr(d_ann, z["d_ann"]) < 2e-5


@pytest.mark.parametrize("k", [1, 3, 5])
def test_decode_vs_reference_golden(k):
    """two layers: token ids bit-exact, scores / alphas at 1e-5"""
    z, W, _ = load_golden("decode_layers2")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    for rescore in (None, "LN", "WR", "BAR"):
        caps, scores, alphas, ppl = cuda_caption(W, ann, k, max_len, 1.0, rescore, 0.5, True)
        tag = "k%d_%s_all" % (k, rescore)
        for i in range(ann.shape[0]):
            assert len(caps[i]) == int(z["%s/n%d/count" % (tag, i)])
            for j in range(len(caps[i])):
                assert caps[i][j] == z["%s/n%d/h%d/tokens" % (tag, i, j)].tolist()
                ref_s = float(z["%s/n%d/h%d/score" % (tag, i, j)])
                assert abs(scores[i][j] - ref_s) < 1e-5 * max(1.0, abs(ref_s))
                assert relerr(alphas[i][j], z["%s/n%d/h%d/alphas" % (tag, i, j)]) < 1e-5


@pytest.mark.parametrize("layers", [2, 4])
def test_c1_dims_fp32_vs_oracle(layers):
    cfg = dict(Bi=6, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=1000, T=7, ragged=True, layers=layers)
    W, ann, caps, lens = synth(**cfg)
    W["attention.f_att.weight"] *= 10
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.1, 1.0)
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, 0.1, 1.0)
    assert abs(loss - loss_ref) < 1e-5 * abs(loss_ref)
    for k, g in Gref.items():
        assert relerr(G[k], g) < 5e-5, k
    assert relerr(d_ann, da_ref) < 5e-5


@pytest.mark.parametrize("use_tc", [False, True])
def test_c1_dims_bf16_vs_oracle(use_tc):
    """bf16 (tensor-core GEMMs with split-K on the per-layer chains) at BASELINE configs[0] decoder dims, two layers"""
    cfg = dict(Bi=8, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=10, ragged=True, layers=2)
    W, ann, caps, lens = synth(**cfg)
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.0, 1.0)
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=use_tc)
    assert abs(loss - loss_ref) < 2e-2 * abs(loss_ref)
    for k, g in Gref.items():
        assert relerr(G[k], g) < 6e-2, k
    assert relerr(d_ann, da_ref) < 6e-2


def test_bf16_gradients_are_bit_reproducible():
    cfg = dict(Bi=8, ncap=2, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=True, layers=2)
    W, ann, caps, lens = synth(**cfg)
    _, G1, da1 = run_cuda_fwd_bwd(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=True)
    _, G2, da2 = run_cuda_fwd_bwd(W, ann, caps, lens, 0.0, 1.0, dtype=torch.bfloat16, exact=False, use_tc=True)
    for k in G1:
        assert torch.equal(G1[k], G2[k]), k
    assert torch.equal(da1, da2)


def test_decode_c1_dims_vs_oracle():
    cfg = dict(Bi=4, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=1000, T=4, ragged=False, layers=3)
    W, ann, _, _ = synth(**cfg, sharpen=True)
    for k in (1, 3):
        ref = O.caption(W, ann, VOC(1000), beamk=k, max_gen_length=10, rescore_method="LN", return_all=False)
        got = cuda_caption(W, ann, k, 10, 1.0, "LN", 0.5, False)
        assert got[0] == ref[0]
        for a, b in zip(got[1], ref[1]):
            assert abs(a - b) < 1e-4 * max(1.0, abs(b))


def _module(layers, **over):
    from sat_b200.model import SAT
    torch.manual_seed(0)
    hp = rh.default_hparams(encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128, input_size=64,
                            decoder_layers=layers, **over)
    m = SAT(**hp)
    m.encoder = nn.Identity()
    with torch.no_grad():
        m.attention.f_att.weight *= 6
    return m.cuda()


def test_module_training_step_and_caption():
    """SAT(decoder_layers=2): training_step loss / every gradient (incl. lstm.*_l1) against the oracle; caption() against the
    oracle's beam search; decoder noise touches every layer's state"""
    m = _module(2, label_smoothing=0.05)
    g = torch.Generator().manual_seed(2)
    ann = torch.randn(5, 64, 4, 3, generator=g)
    caps = torch.randint(1, 125, (5, 2, 8), generator=g)
    caps[:, :, 0] = 126
    lens = torch.randint(2, 8, (5, 2), generator=g)
    W = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items() if not k.startswith("encoder")}
    assert "lstm.weight_ih_l1" in W and W["init_lstm.init.weight"].shape[0] == 2 * 2 * 64
    ref = O.train_loss(W, ann, caps, lens, 0.05, 1.0)
    ref["loss"].backward()
    m.train()
    out = m.training_step((ann.cuda(), caps.cuda(), lens.cuda()), 0)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    out["loss"].backward()
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    m.eval()
    Wd = {k: v.detach() for k, v in W.items()}
    with torch.no_grad():
        m.output.output.weight *= 6
        Wd["output.output.weight"] = Wd["output.output.weight"] * 6
    ref_caps = O.caption(Wd, ann, VOC(128), beamk=3, max_gen_length=8, rescore_method="LN")[0]
    got = m.caption(ann.cuda(), beamk=3, max_gen_length=8, rescore_method="LN")[0]
    assert got == ref_caps
    a = m.caption(ann.cuda(), beamk=1, max_gen_length=8, decoder_noise=0.0)[0]
    torch.manual_seed(3)
    b = m.caption(ann.cuda(), beamk=1, max_gen_length=8, decoder_noise=50.0)[0]
    assert a == m.caption(ann.cuda(), beamk=1, max_gen_length=8)[0]
    assert a != b                                    # a large noise on the recurrent inputs changes the greedy captions
