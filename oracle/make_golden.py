"""Generate tests/golden/*.npz from the UNMODIFIED reference (run by hand in the build container:
`python oracle/make_golden.py`).  Test infrastructure only; needs /root/reference, which does not
exist on the GPU box -- the committed .npz files are what travels.

Every case drives the reference's own public methods (SAT.train_batch, SAT.criterion, the
doubly-stochastic term exactly as training_step writes it, SAT.caption) with the CNN trunk replaced
by nn.Identity() so that the "image" tensor IS the annotation tensor [B,D,h,w]; decoder weights
are the reference's own default init under the stated seed (plus the stated sharpening).
"""
import os
import sys
import warnings

import numpy as np
import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

warnings.filterwarnings("ignore")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build(model, seed, D, A, E, H, V, label_smoothing=0.0, att_gamma=1.0, sharpen=None):
    torch.manual_seed(seed)
    hp = rh.default_hparams(encoder_dim=D, attention_dim=A, embed_dim=E, decoder_dim=H, vocab_size=V,
                            label_smoothing=label_smoothing, att_gamma=att_gamma, input_size=64)
    m = model.SAT(**hp)
    m.encoder = nn.Identity()
    if sharpen:
        with torch.no_grad():
            m.output.output.weight *= sharpen.get("wo", 1.0)
            m.embedding.weight *= sharpen.get("emb", 1.0)
            m.attention.f_att.weight *= sharpen.get("fatt", 1.0)
            if "end_bias" in sharpen:
                m.output.output.bias[V - 1] = sharpen["end_bias"]
    return m


def weights_of(m):
    return {"W/" + k: v.detach().numpy().copy() for k, v in m.state_dict().items() if not k.startswith("encoder")}


def train_case(model, name, seed, B_img, ncap, hw, D, A, E, H, V, T, ragged, label_smoothing, sharpen=None):
    m = build(model, seed, D, A, E, H, V, label_smoothing, 1.0, sharpen)
    g = torch.Generator().manual_seed(seed + 1)
    ann = torch.randn(B_img, D, hw[0], hw[1], generator=g)
    ann.requires_grad_(True)
    caps = torch.randint(1, V - 3, (B_img, ncap, T + 1), generator=g)
    caps[:, :, 0] = V - 2
    if ragged:
        lengths = torch.randint(2, T + 1, (B_img, ncap), generator=g)
    else:
        lengths = torch.full((B_img, ncap), T)
    for i in range(B_img):
        for j in range(ncap):
            n = int(lengths[i, j])
            caps[i, j, n] = V - 1            # <END> is the last target
            caps[i, j, n + 1:] = 0           # <PAD>
    m.train()
    lp, tp, alphas = m.train_batch([ann, caps, lengths], torch.tensor(1))
    ce = m.criterion(lp.data, tp.data)
    loss = ce + m.hparams.att_gamma * ((1 - alphas.sum(dim=1)) ** 2).mean()      # model.py:592-594
    pred = torch.argmax(lp.data, dim=1)
    acc = torch.sum(pred == tp.data) / pred.shape[0]
    loss.backward()
    out = weights_of(m)
    for k, p in m.named_parameters():
        if not k.startswith("encoder") and p.grad is not None:
            out["G/" + k] = p.grad.numpy().copy()
    padded = torch.zeros(B_img * ncap, T, V)
    # unpack to padded [B,T,V] for convenience (time-major packed data is also stored)
    from torch.nn.utils.rnn import pad_packed_sequence
    padded, _ = pad_packed_sequence(lp, batch_first=True, total_length=T)
    out.update(dict(ann=ann.detach().numpy(), caps=caps.numpy(), lengths=lengths.numpy(),
                    logits_packed=lp.data.detach().numpy(), targets_packed=tp.data.numpy(),
                    logits=padded.detach().numpy(), alphas=alphas.detach().numpy(),
                    loss=np.float64(loss.item()), ce=np.float64(ce.item()), acc=np.float64(acc.item()),
                    d_ann=ann.grad.numpy().copy(), label_smoothing=np.float64(label_smoothing),
                    att_gamma=np.float64(1.0), dims=np.array([D, A, E, H, V, T, hw[0], hw[1], ncap])))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "acc", acc.item(), "tokens", lp.data.shape[0])


def decode_case(model, name, seed, n_img, hw, D, A, E, H, V, max_len, sharpen):
    m = build(model, seed, D, A, E, H, V, 0.0, 1.0, sharpen)
    g = torch.Generator().manual_seed(seed + 1)
    ann = torch.randn(n_img, D, hw[0], hw[1], generator=g)
    out = weights_of(m)
    out["ann"] = ann.numpy()
    out["dims"] = np.array([D, A, E, H, V, max_len, hw[0], hw[1]])
    lens = {}
    for k in (1, 3, 5):
        for rs in (None, "LN", "WR", "BAR"):
            for ra in (False, True):
                caps, scores, alphas, ppl = m.caption(ann.clone(), beamk=k, max_gen_length=max_len, temperature=1.0,
                                                      rescore_method=rs, rescore_reward=0.5, return_all=ra)
                tag = "k%d_%s_%s" % (k, rs, "all" if ra else "best")
                if not ra:
                    caps, scores, alphas, ppl = [[c] for c in caps], [[s] for s in scores], [[a] for a in alphas], [[p] for p in ppl]
                lens[tag] = [[len(c) for c in cc] for cc in caps]
                for i in range(n_img):
                    out["%s/n%d/count" % (tag, i)] = np.array(len(caps[i]))
                    for j in range(len(caps[i])):
                        out["%s/n%d/h%d/tokens" % (tag, i, j)] = np.array(caps[i][j], dtype=np.int64)
                        out["%s/n%d/h%d/score" % (tag, i, j)] = np.float64(scores[i][j])
                        out["%s/n%d/h%d/ppl" % (tag, i, j)] = np.float64(ppl[i][j])
                        out["%s/n%d/h%d/alphas" % (tag, i, j)] = alphas[i][j].numpy()
    # temperature != 1
    caps, scores, alphas, ppl = m.caption(ann.clone(), beamk=3, max_gen_length=max_len, temperature=0.7,
                                          rescore_method="LN", return_all=False)
    for i in range(n_img):
        out["k3_LN_T0.7/n%d/tokens" % i] = np.array(caps[i], dtype=np.int64)
        out["k3_LN_T0.7/n%d/score" % i] = np.float64(scores[i])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v for k, v in lens.items() if k.endswith("None_best") or k.endswith("LN_best")})


def main():
    os.makedirs(OUT, exist_ok=True)
    model, _ = rh.load_reference()
    # tiny, ragged, 2 captions per image, non-square map, label smoothing, peaky attention
    train_case(model, "train_tiny", 0, B_img=3, ncap=2, hw=(3, 4), D=16, A=8, E=10, H=14, V=50, T=6,
               ragged=True, label_smoothing=0.1, sharpen=dict(fatt=20.0))
    # small, square map, fixed length (BASELINE-style: ncap 1, all lengths T), plain CE
    train_case(model, "train_small", 1, B_img=5, ncap=1, hw=(7, 7), D=64, A=32, E=32, H=48, V=120, T=8,
               ragged=False, label_smoothing=0.0)
    # small ragged with dims that are multiples of the CUDA tile sizes
    train_case(model, "train_ragged", 2, B_img=4, ncap=1, hw=(4, 4), D=64, A=32, E=32, H=64, V=128, T=10,
               ragged=True, label_smoothing=0.05, sharpen=dict(fatt=10.0))
    # decode: sharpened so that <END> is reachable and beams shrink
    decode_case(model, "decode_tiny", 3, n_img=6, hw=(3, 4), D=16, A=8, E=10, H=14, V=50, max_len=12,
                sharpen=dict(wo=8.0, emb=2.0, fatt=30.0, end_bias=2.0))
    decode_case(model, "decode_small", 4, n_img=5, hw=(4, 4), D=64, A=32, E=32, H=64, V=128, max_len=16,
                sharpen=dict(wo=8.0, emb=2.0, fatt=30.0, end_bias=3.0))


if __name__ == "__main__":
    main()
