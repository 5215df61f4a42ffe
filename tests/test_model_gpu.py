"""GPU: the SAT module API (train_batch / fused training loss / autograd) against the CPU oracle, with the CNN
trunk replaced by nn.Identity() so that decoder parity is measured on identical annotations."""
import warnings

import pytest
import torch
from torch import nn

from oracle import ref_harness as rh
from oracle import sat_oracle as O
from test_train_forward_gpu import relerr

warnings.filterwarnings("ignore")
pytestmark = pytest.mark.gpu


def build(seed=0, **over):
    from sat_b200.model import SAT
    torch.manual_seed(seed)
    hp = rh.default_hparams(encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128, input_size=64, **over)
    m = SAT(**hp)
    m.encoder = nn.Identity()
    with torch.no_grad():
        m.attention.f_att.weight *= 6        # non-uniform attention (softmax backward is exercised).  The scores' rounding error is
    return m.cuda()                          # amplified by this factor: x10 sat on the 1e-5 tolerance (1.8e-5 on one GPU host, whose
                                             # CPU runs the fp32 oracle's GEMMs in a different order)


def batch(seed, Bi=5, ncap=2, T=7, V=128, D=64, hw=(4, 3)):
    g = torch.Generator().manual_seed(seed)
    ann = torch.randn(Bi, D, hw[0], hw[1], generator=g)
    caps = torch.randint(1, V - 3, (Bi, ncap, T + 1), generator=g)
    caps[:, :, 0] = V - 2
    lens = torch.randint(2, T + 1, (Bi, ncap), generator=g)
    return ann, caps, lens


def weights_cpu(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if not k.startswith("encoder")}


def test_train_batch_matches_oracle_and_is_differentiable():
    m = build(label_smoothing=0.1)
    ann, caps, lens = batch(1)
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    a_ref = ann.clone().requires_grad_(True)
    ref = O.train_loss(W, a_ref, caps, lens, 0.1, 1.0)
    ref["loss"].backward()
    m.train()
    a = ann.cuda().requires_grad_(True)
    lp, tp, alphas = m.train_batch((a, caps.cuda(), lens.cuda()), epsilon=1)
    assert lp.data.dtype == torch.float32                                  # model.py:504
    assert torch.equal(tp.data.cpu(), ref["targets_packed"].data)
    assert torch.equal(lp.batch_sizes, ref["logits_packed"].batch_sizes)
    assert relerr(lp.data, ref["logits_packed"].data) < 1e-5
    assert relerr(alphas, ref["alphas"]) < 1e-5
    loss = m.criterion(lp.data, tp.data) + m.hparams.att_gamma * ((1 - alphas.sum(dim=1)) ** 2).mean()   # model.py:592-594
    assert abs(float(loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    loss.backward()
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    assert relerr(a.grad, a_ref.grad) < 5e-5


def test_training_step_fused_loss_and_grads():
    m = build(seed=1, label_smoothing=0.05)
    ann, caps, lens = batch(2, ncap=1)
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    ref = O.train_loss(W, ann, caps, lens, 0.05, 1.0)
    ref["loss"].backward()
    m.train()
    out = m.training_step((ann.cuda(), caps.cuda(), lens.cuda()), 0)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    assert abs(float(out["accuracy"]) - float(ref["acc"])) < 1e-6
    out["loss"].backward()
    for k, p in m.named_parameters():
        assert relerr(p.grad, W[k].grad) < 5e-5, k


def test_module_forwards_match_oracle():
    m = build(seed=2)
    ann, _, _ = batch(3, ncap=1)
    W = weights_cpu(m)
    h = torch.randn(5, 64)
    z_ref, al_ref = O.attention(W, ann, h)
    z, al = m.attention(ann.cuda(), h.cuda())
    assert al.shape == (5, 4, 3)
    assert relerr(z, z_ref) < 1e-5 and relerr(al.reshape(5, -1), al_ref) < 1e-5
    h0_ref, c0_ref = O.init_lstm(W, ann)
    h0, c0 = m.init_lstm(ann.cuda())
    assert h0.shape == (1, 5, 64)
    assert relerr(h0[0], h0_ref) < 1e-5 and relerr(c0[0], c0_ref) < 1e-5
    xe = torch.randn(5, 32)
    lo_ref = O.deep_output(W, xe, h, z_ref)
    lo = m.output(xe.cuda(), h.cuda(), z_ref.cuda())
    assert relerr(lo, lo_ref) < 1e-5


@pytest.mark.parametrize("epsilon", [0.0, 0.5])
def test_scheduled_sampling_matches_oracle(epsilon):
    """epsilon < 1 (model.py:518-523): steps > 2 feed back argmax(logits[t-1]) when the host draw exceeds epsilon.  Same CPU RNG
    stream for the oracle and the module (one torch.rand(1) per step > 2 that still has an active caption)."""
    m = build(seed=3, label_smoothing=0.1)
    with torch.no_grad():
        m.output.output.weight *= 6          # decisive argmax so the fed-back tokens are stable across rounding
    ann, caps, lens = batch(4, ncap=2, T=9)
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    a_ref = ann.clone().requires_grad_(True)
    torch.manual_seed(123)
    ref = O.train_loss(W, a_ref, caps, lens, 0.1, 1.0, epsilon=epsilon)
    ref["loss"].backward()
    m.train()
    torch.manual_seed(123)
    a = ann.cuda().requires_grad_(True)
    lp, tp, alphas = m.train_batch((a, caps.cuda(), lens.cuda()), epsilon=epsilon)
    assert relerr(lp.data, ref["logits_packed"].data) < 1e-5
    assert relerr(alphas, ref["alphas"]) < 1e-5
    loss = m.criterion(lp.data, tp.data) + ((1 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    for k, p in m.named_parameters():
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    # fused path with the same draws
    m.zero_grad()
    torch.manual_seed(123)
    loss2, aux = m.fused_loss((ann.cuda(), caps.cuda(), lens.cuda()), epsilon=epsilon)
    assert abs(float(loss2) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    loss2.backward()
    assert relerr(m.embedding.weight.grad, W["embedding.weight"].grad) < 5e-5


def test_weight_tying_shares_embedding_and_grads_match():
    """weight_tying (model.py:198-199): output.output.weight IS embedding.weight, no output bias."""
    from sat_b200.model import SAT
    torch.manual_seed(5)
    hp = rh.default_hparams(encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128, input_size=64,
                            weight_tying=True, label_smoothing=0.1)
    m = SAT(**hp)
    m.encoder = nn.Identity()
    m = m.cuda()
    assert m.output.output.weight is m.embedding.weight and m.output.output.bias is None
    ann, caps, lens = batch(7, ncap=1)
    W = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items() if not k.startswith("encoder")}
    emb = W["embedding.weight"]
    Wt = dict(W)
    Wt["output.output.weight"] = emb                      # tied: same leaf
    Wt["output.output.bias"] = None
    ref = O.train_loss(Wt, ann, caps, lens, 0.1, 1.0)
    ref["loss"].backward()
    m.train()
    out = m.training_step((ann.cuda(), caps.cuda(), lens.cuda()), 0)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    out["loss"].backward()
    assert relerr(m.embedding.weight.grad, emb.grad) < 5e-5
    assert relerr(m.lstm.weight_ih_l0.grad, W["lstm.weight_ih_l0"].grad) < 5e-5


def test_plain_output_layer_deep_output_false():
    """deep_output=False (train.py default; model.py:128-129): x = W_ho h', no tanh / embedding / context term."""
    m = build(seed=6, deep_output=False, label_smoothing=0.1)
    assert not hasattr(m.output, "context")
    ann, caps, lens = batch(8, ncap=1)
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    a_ref = ann.clone().requires_grad_(True)
    ref = O.train_loss(W, a_ref, caps, lens, 0.1, 1.0, deep=False)
    ref["loss"].backward()
    m.train()
    a = ann.cuda().requires_grad_(True)
    out = m.training_step((a, caps.cuda(), lens.cuda()), 0)
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    out["loss"].backward()
    for k, p in m.named_parameters():
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    assert relerr(a.grad, a_ref.grad) < 5e-5
    # decode through the module API
    Wd = {k: v.detach() for k, v in W.items()}
    Wd["output.output.weight"] = Wd["output.output.weight"] * 8
    with torch.no_grad():
        m.output.output.weight *= 8
    vocab = dict(PAD=0, UNK=125, START=126, END=127)
    ref_c = O.caption(Wd, ann, vocab, beamk=3, max_gen_length=8, rescore_method="LN", deep=False)
    got = m.caption(ann.cuda(), beamk=3, max_gen_length=8, rescore_method="LN")
    assert got[0] == ref_c[0]
    assert max(abs(x - y) for x, y in zip(got[1], ref_c[1])) < 1e-4


def test_caption_api_returns_reference_format():
    m = build(seed=7)
    ann, _, _ = batch(9, ncap=1)
    caps, scores, alphas, ppl = m.caption(ann.cuda(), beamk=2, max_gen_length=5, return_all=True)
    assert not m.training                                        # caption() leaves the module in eval() (model.py:231)
    assert len(caps) == ann.shape[0] and all(isinstance(c, list) and isinstance(c[0], list) for c in caps)
    assert all(a[0].device.type == "cpu" and a[0].shape[1:] == (4, 3) for a in alphas)
    assert all(len(c[0]) == a[0].shape[0] for c, a in zip(caps, alphas))
    assert all(s == sorted(s, reverse=True) for s in scores)


def test_dropout_train_mode_matches_oracle_with_the_same_masks():
    """dropout / embedding_dropout > 0 in train() mode (model.py:78,526,130; the reference's recommended recipes use 0.2):
    the kernels' masks are a pure function of (seed, index), so the oracle is run with exactly those masks."""
    import ctypes as C
    from sat_b200 import _lib
    m = build(seed=9, label_smoothing=0.1, dropout=0.3, embedding_dropout=0.2)
    ann, caps, lens = batch(11, ncap=1, T=6)
    Bi, D, E, T = ann.shape[0], 64, 32, 6
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    m.train()
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # the draw fused_loss will make
    mul = lambda p, stream, idx: _lib.lib().sat_dropout_multiplier(C.c_float(p), C.c_uint64(seed), C.c_uint32(stream), C.c_uint64(idx))
    mean_mask = torch.tensor([[mul(0.3, 1, i * D + d) for d in range(D)] for i in range(Bi)])
    emb_mask = torch.tensor([[[mul(0.2, 2, (t * Bi + b) * E + e) for e in range(E)] for t in range(T)] for b in range(Bi)])
    out_mask = torch.tensor([[[mul(0.3, 3, (t * Bi + b) * E + e) for e in range(E)] for t in range(T)] for b in range(Bi)])
    assert 0.15 < float((emb_mask == 0).float().mean()) < 0.25 and 0.22 < float((out_mask == 0).float().mean()) < 0.38
    a_ref = ann.clone().requires_grad_(True)
    ref = O.train_loss(W, a_ref, caps, lens, 0.1, 1.0, masks=dict(mean=mean_mask, emb=emb_mask, out=out_mask))
    ref["loss"].backward()
    torch.manual_seed(77)
    a = ann.cuda().requires_grad_(True)
    loss, aux = m.fused_loss((a, caps.cuda(), lens.cuda()))
    assert abs(float(loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    loss.backward()
    for k, p in m.named_parameters():
        assert relerr(p.grad, W[k].grad) < 5e-5, k
    assert relerr(a.grad, a_ref.grad) < 5e-5
    # eval mode ignores dropout
    m.eval()
    l_eval, _ = m.fused_loss((ann.cuda(), caps.cuda(), lens.cuda()))
    ref0 = O.train_loss({k: v.detach() for k, v in W.items()}, ann, caps, lens, 0.1, 1.0)
    assert abs(float(l_eval) - float(ref0["loss"])) < 1e-5 * abs(float(ref0["loss"]))


def test_embed_norm_renormalises_like_nn_embedding():
    """embed_norm (nn.Embedding(max_norm=...), model.py:161): looked-up rows are clipped in place before use."""
    m = build(seed=10, embed_norm=1.0)
    ann, caps, lens = batch(12, ncap=1)
    W = weights_cpu(m)
    ref_emb = torch.nn.Embedding(128, 32, max_norm=1.0, padding_idx=0)
    with torch.no_grad():
        ref_emb.weight.copy_(W["embedding.weight"])
    ref_emb(caps[..., :-1].reshape(-1))                       # in-place renorm of the used rows, as the reference's forward does
    Wr = dict(W)
    Wr["embedding.weight"] = ref_emb.weight.detach().clone()
    ref = O.train_loss(Wr, ann, caps, lens, 0.0, 1.0)
    m.train()
    loss, _ = m.fused_loss((ann.cuda(), caps.cuda(), lens.cuda()))
    assert abs(float(loss) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    assert relerr(m.embedding.weight, ref_emb.weight) < 1e-6      # same side effect on the parameter
    used = caps[..., :-1].reshape(-1).unique()
    assert float(m.embedding.weight[used.cuda()].norm(dim=1).max()) <= 1.0 + 1e-5


def test_caption_stream_matches_caption():
    """bulk captioning over pinned host batches returns, batch by batch, exactly what caption() returns"""
    m = build(seed=11)
    batches = [batch(20 + i, ncap=1)[0].pin_memory() for i in range(3)]       # encoder = Identity: "images" are annotations
    got = list(m.caption_stream(batches, beamk=3, max_gen_length=8, rescore_method="LN"))
    assert len(got) == 3
    for b, out in zip(batches, got):
        ref = m.caption(b.cuda(), beamk=3, max_gen_length=8, rescore_method="LN")
        assert out[0] == ref[0]
        assert max(abs(x - y) for x, y in zip(out[1], ref[1])) < 1e-6


@pytest.mark.parametrize("layers", [1, 2])
def test_training_loop_trajectory_matches_oracle_with_adam(layers):
    """End-to-end training loop through the module API (training_step -> backward -> Adam -> repack of the kernel weights, 6
    steps on one batch, fp32): the loss trajectory equals the CPU oracle trained with the same torch.optim.Adam; afterwards the
    bf16 path overfits the same batch (loss falls by more than half in 60 steps)."""
    m = build(seed=7, label_smoothing=0.0, decoder_layers=layers, decoder_lr=2e-3, embedding_lr=2e-3)
    ann, caps, lens = batch(8, ncap=2)
    W = {k: v.requires_grad_(True) for k, v in weights_cpu(m).items()}
    ref_opt = torch.optim.Adam(list(W.values()), lr=2e-3)
    opt = m.configure_optimizers()
    m.train()
    for step in range(6):
        ref = O.train_loss(W, ann, caps, lens, 0.0, 1.0)
        ref_opt.zero_grad()
        ref["loss"].backward()
        ref_opt.step()
        out = m.training_step((ann.cuda(), caps.cuda(), lens.cuda()), step)
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        assert abs(float(out["loss"]) - float(ref["loss"])) < 2e-4 * abs(float(ref["loss"])), step
    assert float(out["loss"]) < float(O.train_loss(weights_cpu(build(seed=7, decoder_layers=layers)), ann, caps, lens, 0.0, 1.0)["loss"])
    # bf16 / tensor cores: overfit one batch
    mb = build(seed=7, label_smoothing=0.0, decoder_layers=layers, decoder_lr=2e-3, embedding_lr=2e-3, precision="bf16")
    optb = mb.configure_optimizers()
    mb.train()
    first = None
    for step in range(60):
        out = mb.training_step((ann.cuda(), caps.cuda(), lens.cuda()), step)
        optb.zero_grad(set_to_none=True)
        out["loss"].backward()
        optb.step()
        first = float(out["loss"]) if first is None else first
    assert float(out["loss"]) < 0.5 * first, (first, float(out["loss"]))


def test_val_batch_and_validation_step_against_the_reference_callers():
    """§8 f4: val_batch / validation_step (model.py:684-697) = caption() + score_captions().  Ours on the GPU against the unmodified
    reference on the CPU with the same weights and batch (when the staged reference is present; its nltk imports are pointed at
    sat_b200.metrics, whose BLEU / GLEU are pinned to nltk's docstring values in tests/test_callers_cpu.py): identical captions,
    hence identical scores."""
    import sys
    m = build(seed=11)
    with torch.no_grad():
        m.output.output.weight *= 6
    g = torch.Generator().manual_seed(12)
    ann = torch.randn(4, 64, 4, 3, generator=g)
    caps = torch.randint(1, 124, (4, 3, 9), generator=g)
    caps[:, :, 0] = 126
    lens = torch.randint(2, 8, (4, 3), generator=g)
    for i in range(4):
        for j in range(3):
            caps[i, j, int(lens[i, j])] = 127
            caps[i, j, int(lens[i, j]) + 1:] = 0
    out = m.val_batch((ann.cuda(), caps.cuda(), lens.cuda()), beamk=3, max_gen_length=8, temperature=1.0, rescore_method="LN")
    assert {"bleu1", "bleu2", "bleu3", "bleu4", "gleu", "cosine_similarity", "perplexity"} <= set(out.keys())
    got_caps, _, _, ppl = m.caption(ann.cuda(), beamk=3, max_gen_length=8, temperature=1.0, rescore_method="LN")
    again = m.score_captions(got_caps, caps, lens, ppl)
    assert all(abs(out[k] - again[k]) < 1e-9 for k in out)
    m.hparams["val_beamk"], m.hparams["val_max_len"] = 3, 8
    vs = m.validation_step((ann.cuda(), caps.cuda(), lens.cuda()), 0)
    assert all(abs(out[k] - vs[k]) < 1e-9 for k in out)
    if not rh.available():
        return
    model_mod, _ = rh.load_reference()
    from sat_b200 import metrics
    model_mod.corpus_bleu, model_mod.corpus_gleu = metrics.corpus_bleu, metrics.corpus_gleu
    hp = rh.default_hparams(encoder_dim=64, attention_dim=32, embed_dim=32, decoder_dim=64, vocab_size=128, input_size=64)
    ref = model_mod.SAT(**hp)
    ref.encoder = nn.Identity()
    ref.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    want = ref.val_batch([ann.clone(), caps, lens], beamk=3, max_gen_length=8, temperature=1.0, rescore_method="LN")
    assert set(want.keys()) == set(out.keys())
    for k in want:
        assert abs(float(want[k]) - float(out[k])) < 1e-5 * max(1.0, abs(float(want[k]))), k
