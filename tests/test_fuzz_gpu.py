"""GPU: seeded random configurations of the training path (dims that are / are not multiples of 8, 1-3 stacked LSTM layers, 1-5
captions per image, deep / plain output layer, with / without an output bias, ragged lengths, label smoothing) in fp32 against
autograd of the CPU oracle, then decoded greedily and with a beam against the oracle's token ids.  Complements the hand-picked
cases: every combination goes through the same drivers in one process, in sequence (shared kernels, caches, attribute state)."""
import random

import pytest
import torch

from oracle import sat_oracle as O
from test_decode_gpu import VOC, cuda_caption
from test_train_backward_gpu import run_cuda_fwd_bwd
from test_train_forward_gpu import relerr

pytestmark = pytest.mark.gpu


def _case(seed):
    r = random.Random(seed)
    mult8 = r.random() < 0.4
    dim = (lambda lo, hi: r.randrange(lo, hi, 8)) if mult8 else (lambda lo, hi: r.randrange(lo, hi))
    cfg = dict(D=dim(16, 97), A=dim(8, 49), E=dim(8, 65), H=dim(8, 73), V=r.randrange(24, 200) if not mult8 else r.randrange(24, 200, 8),
               hw=(r.randrange(1, 6), r.randrange(1, 6)), Bi=r.randrange(1, 6), ncap=r.choice([1, 1, 2, 3, 5]), T=r.randrange(2, 9),
               layers=r.choice([1, 1, 2, 3]), deep=r.random() < 0.7, bias=r.random() < 0.7, ls=r.choice([0.0, 0.1]))
    return cfg


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_config_fp32_vs_oracle(seed):
    c = _case(seed)
    W = O.random_weights(c["D"], c["A"], c["E"], c["H"], c["V"], seed=seed, layers=c["layers"], sharpen=True)
    if not c["deep"]:
        del W["output.context.weight"]
    if not c["bias"]:
        del W["output.output.bias"]
    g = torch.Generator().manual_seed(1000 + seed)
    ann = torch.randn(c["Bi"], c["D"], *c["hw"], generator=g)
    V, T = c["V"], c["T"]
    caps = torch.randint(1, V - 3, (c["Bi"], c["ncap"], T + 1), generator=g)
    caps[:, :, 0] = V - 2
    lens = torch.randint(1, T + 1, (c["Bi"], c["ncap"]), generator=g)
    Wg = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    a = ann.clone().requires_grad_(True)
    ref = O.train_loss(Wg, a, caps, lens, c["ls"], 1.0, deep=c["deep"])
    ref["loss"].backward()
    loss, G, d_ann = run_cuda_fwd_bwd(W, ann, caps, lens, c["ls"], 1.0)
    assert abs(loss - float(ref["loss"])) < 2e-5 * abs(float(ref["loss"])), c
    assert set(G.keys()) == set(Wg.keys()), c
    for k, v in Wg.items():
        assert tuple(G[k].shape) == tuple(v.shape), (k, c)
        assert relerr(G[k], v.grad) < 1e-4, (k, c)
    assert relerr(d_ann, a.grad) < 1e-4, c
    # decode with the same weights: greedy and beam 3, token ids bit-exact
    Wd = {k: v for k, v in W.items()}
    for k in (1, 3):
        want = O.caption(Wd, ann, VOC(V), beamk=k, max_gen_length=7, rescore_method="LN", deep=c["deep"])
        got = cuda_caption(Wd, ann, k, 7, 1.0, "LN", 0.5, False)
        assert got[0] == want[0], (k, c)


@pytest.mark.parametrize("seed", list(range(100, 110)))
def test_random_config_bf16_tensor_core_vs_oracle(seed):
    """same generator, bf16 operands on the tcgen05 path (fused vocabulary cross entropy, NT weight-gradient GEMMs, zero-padded
    storage for the odd sizes): loss to 2e-2, gradients to 8e-2 of their largest entry"""
    from sat_b200 import decoder
    from sat_b200.packing import PackedWeights
    c = _case(seed)
    W = O.random_weights(c["D"], c["A"], c["E"], c["H"], c["V"], seed=seed, layers=c["layers"])
    if not c["deep"]:
        del W["output.context.weight"]
    if not c["bias"]:
        del W["output.output.bias"]
    g = torch.Generator().manual_seed(1000 + seed)
    ann = torch.randn(c["Bi"], c["D"], *c["hw"], generator=g)
    V, T = c["V"], c["T"]
    caps = torch.randint(1, V - 3, (c["Bi"], c["ncap"], T + 1), generator=g)
    caps[:, :, 0] = V - 2
    lens = torch.randint(1, T + 1, (c["Bi"], c["ncap"]), generator=g)
    Wg = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    a = ann.clone().requires_grad_(True)
    ref = O.train_loss(Wg, a, caps, lens, c["ls"], 1.0, deep=c["deep"])
    ref["loss"].backward()
    pw = PackedWeights(W, dtype=torch.bfloat16, device="cuda")
    bld = decoder.annotations_as_bld(ann.cuda(), torch.bfloat16)
    buf = decoder.train_forward(pw, bld, caps.cuda(), lens.cuda(), c["ls"], 1.0, exact=False, use_tc=True, backward=True, fuse_ce=True)
    G, d_ann = decoder.train_backward(pw, buf)
    torch.cuda.synchronize()
    loss = float(buf.t["out"][0])
    assert abs(loss - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"])), c
    for k, v in Wg.items():
        assert torch.isfinite(G[k]).all() and relerr(G[k], v.grad) < 8e-2, (k, c)
    Bi, D, h, w = ann.shape
    assert relerr(d_ann.float().reshape(Bi, h, w, D).permute(0, 3, 1, 2), a.grad) < 8e-2, c
