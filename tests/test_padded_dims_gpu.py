"""GPU: modules whose sizes are not multiples of 8 (the reference has no such limit: V = words above min_count + 4,
100/300-d GloVe embeddings).  PackedWeights zero-pads to the kernels' storage dims; forward, loss, every gradient and the
decoded token ids must equal the reference goldens / the CPU oracle on the TRUE dims."""
import pytest
import torch

from conftest import load_golden
from oracle import sat_oracle as O
from test_decode_gpu import VOC, cuda_caption
from test_train_backward_gpu import oracle_grads, run_cuda_fwd_bwd
from test_train_forward_gpu import relerr, run_cuda_forward, synth

pytestmark = pytest.mark.gpu


def test_train_tiny_golden_runs_on_cuda():
    """train_tiny: D=16, A=8, E=10, H=14, V=50, two captions per image, non-square map, ragged, label smoothing, peaky
    attention -- generated from the unmodified reference (oracle/make_golden.py)"""
    z, W, Gref = load_golden("train_tiny")
    ann = torch.from_numpy(z["ann"])
    caps, lens = torch.from_numpy(z["caps"]), torch.from_numpy(z["lengths"])
    r = run_cuda_forward(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    V0 = int(z["dims"][4])
    assert relerr(r["alphas"], z["alphas"]) < 1e-5
    assert relerr(r["logits"][..., :V0], z["logits"]) < 1e-5
    assert abs(r["loss"] - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert abs(r["acc"] - float(z["acc"])) < 1e-6
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, float(z["label_smoothing"]), float(z["att_gamma"]))
    assert abs(loss - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    for k, g in Gref.items():
        assert tuple(G[k].shape) == tuple(g.shape), k
        assert relerr(G[k], g) < 2e-5, k
    assert relerr(d_ann, z["d_ann"]) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_odd_dims_vs_oracle(dtype):
    """GloVe-like E=100, a vocabulary that is not a multiple of 8, odd H / A / D"""
    cfg = dict(Bi=5, ncap=2, hw=(4, 5), D=36, A=20, E=100, H=50, V=1003, T=7, ragged=True)
    W, ann, caps, lens = synth(**cfg, seed=3)
    W["attention.f_att.weight"] *= 10
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.1, 1.0)
    fp32 = dtype == torch.float32
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, 0.1, 1.0, dtype=dtype, exact=fp32, use_tc=not fp32)
    tol, gtol = (1e-5, 5e-5) if fp32 else (2e-2, 6e-2)
    assert abs(loss - loss_ref) < tol * abs(loss_ref)
    for k, g in Gref.items():
        assert tuple(G[k].shape) == tuple(g.shape), k
        assert relerr(G[k], g) < gtol, k
    assert relerr(d_ann, da_ref) < gtol


@pytest.mark.parametrize("k", [1, 3, 5])
def test_decode_tiny_golden_on_cuda(k):
    """decode_tiny: D=16, A=8, E=10, H=14, V=50 -- token ids bit-exact against the unmodified reference"""
    z, W, _ = load_golden("decode_tiny")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    for rescore in (None, "LN", "BAR"):
        caps, scores, alphas, ppl = cuda_caption(W, ann, k, max_len, 1.0, rescore, 0.5, True)
        tag = "k%d_%s_all" % (k, rescore)
        for i in range(ann.shape[0]):
            assert len(caps[i]) == int(z["%s/n%d/count" % (tag, i)])
            for j in range(len(caps[i])):
                assert caps[i][j] == z["%s/n%d/h%d/tokens" % (tag, i, j)].tolist()
                ref_s = float(z["%s/n%d/h%d/score" % (tag, i, j)])
                assert abs(scores[i][j] - ref_s) < 1e-5 * max(1.0, abs(ref_s))
                assert relerr(alphas[i][j], z["%s/n%d/h%d/alphas" % (tag, i, j)]) < 1e-5


def test_module_api_with_odd_dims():
    """SAT module with a 300-d embedding and a 1003-word vocabulary: train_batch logits / alphas / gradients have the
    module's own shapes and match the oracle."""
    import warnings
    from torch import nn
    from oracle import ref_harness as rh
    from sat_b200.model import SAT
    warnings.filterwarnings("ignore")
    torch.manual_seed(0)
    hp = rh.default_hparams(encoder_dim=40, attention_dim=24, embed_dim=300, decoder_dim=52, vocab_size=1003, input_size=64,
                            label_smoothing=0.1)
    m = SAT(**hp)
    m.encoder = nn.Identity()
    m = m.cuda().train()
    g = torch.Generator().manual_seed(1)
    ann = torch.randn(4, 40, 3, 3, generator=g)
    caps = torch.randint(1, 1000, (4, 1, 7), generator=g)
    caps[:, :, 0] = 1001
    lens = torch.tensor([[6], [3], [5], [2]])
    W = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items() if not k.startswith("encoder")}
    ref = O.train_loss(W, ann, caps, lens, 0.1, 1.0)
    ref["loss"].backward()
    lp, tp, alphas = m.train_batch((ann.cuda(), caps.cuda(), lens.cuda()), epsilon=1)
    assert lp.data.shape[1] == 1003
    assert relerr(lp.data, ref["logits_packed"].data) < 1e-5
    out = m.training_step((ann.cuda(), caps.cuda(), lens.cuda()), 0)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    out["loss"].backward()
    for k, p in m.named_parameters():
        assert p.grad.shape == p.shape
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    got = m.caption(ann.cuda(), beamk=3, max_gen_length=6, rescore_method="LN")
    vocab = dict(PAD=0, UNK=1000, START=1001, END=1002)
    ref_c = O.caption({k: v.detach() for k, v in W.items()}, ann, vocab, beamk=3, max_gen_length=6, rescore_method="LN")
    assert got[0] == ref_c[0]
