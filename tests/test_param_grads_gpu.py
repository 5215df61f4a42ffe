"""GPU: parameter gradients inside the library (sat_train_param_grads): the "NT" GEMM cores behind dW = dY^T X against
torch, bit-reproducible gradients from run to run (the reference trains with deterministic=True, train.py:271), and the
fused vocabulary-projection + cross-entropy path against the unfused kernels and the CPU oracle."""
import ctypes as C

import pytest
import torch

from oracle import sat_oracle as O
from test_train_backward_gpu import oracle_grads
from test_train_forward_gpu import relerr, synth

pytestmark = pytest.mark.gpu


def linear_nt(A, B, use_tc, splitk=1):
    from sat_b200 import _lib
    K, N1 = A.shape
    N2 = B.shape[1]
    out = torch.empty(splitk, N1, N2, dtype=torch.float32, device=A.device)
    _lib.check(_lib.lib().sat_linear_nt(_lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out), N2, K, N1, N2,
                                        _lib.dtype_code(A.dtype), 1 if use_tc else 0, splitk, _lib.stream_ptr()), "sat_linear_nt")
    torch.cuda.synchronize()
    return out.sum(0)


NT_SHAPES = [(64, 128, 128), (256, 128, 64), (5120, 256, 512), (300, 136, 72), (1000, 2688, 512), (4, 8, 8), (130, 200, 264),
             (50176, 128, 512)]


@pytest.mark.parametrize("K,N1,N2", NT_SHAPES)
@pytest.mark.parametrize("use_tc", [False, True])
@pytest.mark.parametrize("splitk", [1, 4])
def test_linear_nt_bf16(K, N1, N2, use_tc, splitk):
    g = torch.Generator(device="cuda").manual_seed(K + 3 * N1 + 7 * N2)
    A = torch.randn(K, N1, device="cuda", generator=g).to(torch.bfloat16)
    B = (torch.randn(K, N2, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    ref = (A.double().t() @ B.double()).float()
    out = linear_nt(A, B, use_tc, splitk)
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < (2e-4 if K > 10000 else 5e-5), err          # fp32 accumulation of exact bf16 products over K terms


@pytest.mark.parametrize("K,N1,N2", NT_SHAPES[:6])
def test_linear_nt_fp32(K, N1, N2):
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(K, N1, device="cuda", generator=g)
    B = torch.randn(K, N2, device="cuda", generator=g) / K ** 0.5
    ref = (A.double().t() @ B.double()).float()
    out = linear_nt(A, B, False, 2)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5


def test_linear_nt_tc_strided_operands():
    """operands that are column slices of wider buffers (dG = DY[:, A+D:]), as sat_train_param_grads passes them"""
    g = torch.Generator(device="cuda").manual_seed(6)
    big = torch.randn(640, 2688, device="cuda", generator=g).to(torch.bfloat16)
    A = big[:, 640:]
    B = torch.randn(640, 256, device="cuda", generator=g).to(torch.bfloat16)
    ref = A.float().t() @ B.float()
    out = linear_nt(A, B, True, 2)
    assert float((out - ref).abs().max() / ref.abs().max()) < 5e-5


def _fwd_bwd(W, ann, caps, lens, dtype, use_tc, fuse_ce, ls=0.1, weight_tying=False):
    from sat_b200 import decoder
    from sat_b200.packing import PackedWeights
    fp32 = dtype == torch.float32
    pw = PackedWeights(W, dtype=dtype, device="cuda")
    bld = decoder.annotations_as_bld(ann.cuda(), dtype)
    buf = decoder.train_forward(pw, bld, caps.cuda(), lens.cuda(), ls, 1.0, exact=fp32, use_tc=use_tc, backward=True, fuse_ce=fuse_ce)
    G, d_ann = decoder.train_backward(pw, buf, weight_tying=weight_tying)
    torch.cuda.synchronize()
    return buf, {k: v.clone() for k, v in G.items()}, d_ann.clone()


@pytest.mark.parametrize("dtype,use_tc", [(torch.float32, False), (torch.bfloat16, True)])
def test_gradients_are_bit_reproducible(dtype, use_tc):
    """two runs of forward + backward + parameter gradients give identical bits (no floating-point atomics anywhere:
    split-k partials and the embedding segment sum are added in fixed order)"""
    cfg = dict(Bi=24, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=12, ragged=True)
    W, ann, caps, lens = synth(**cfg, seed=9)
    caps[:, :, 1:4] = 7                       # many rows feed the same word: the embedding gradient sums long segments
    _, G1, d1 = _fwd_bwd(W, ann, caps, lens, dtype, use_tc, fuse_ce=use_tc)
    _, G2, d2 = _fwd_bwd(W, ann, caps, lens, dtype, use_tc, fuse_ce=use_tc)
    for k in G1:
        assert torch.equal(G1[k], G2[k]), k
    assert torch.equal(d1, d2)


def test_fused_vocab_ce_matches_unfused_and_oracle():
    """bf16 / tcgen05: vocabulary GEMM with soft-max statistics in its epilogue + recomputed dlogits against the unfused
    GEMM -> ce_rows_kernel path (same weights, same inputs) and against the CPU oracle at the bf16 tolerance."""
    cfg = dict(Bi=12, ncap=1, hw=(14, 14), D=512, A=128, E=256, H=512, V=6400, T=10, ragged=True)
    W, ann, caps, lens = synth(**cfg, seed=13)
    loss_ref, Gref, da_ref = oracle_grads(W, ann, caps, lens, 0.1, 1.0)
    ref = O.train_loss(W, ann, caps, lens, 0.1, 1.0)
    bf, Gf, df = _fwd_bwd(W, ann, caps, lens, torch.bfloat16, True, fuse_ce=True)
    bu, Gu, du = _fwd_bwd(W, ann, caps, lens, torch.bfloat16, True, fuse_ce=False)
    assert "logits" not in bf.t and "ce_stats" in bf.t and "logits" in bu.t
    lf, lu = float(bf.t["out"][0]), float(bu.t["out"][0])
    assert abs(lf - loss_ref) < 2e-2 * abs(loss_ref) and abs(lf - lu) < 5e-3 * abs(lu)
    assert abs(float(bf.t["out"][3]) - float(ref["acc"])) < 0.05          # accuracy (bf16 ties may flip a few arg-maxes)
    # dlogits of the fused path against the unfused one (the unfused path rounds the logits to bf16 first)
    assert relerr(bf.t["dlogits"].float(), bu.t["dlogits"].float()) < 3e-2
    for k, g in Gref.items():
        assert relerr(Gf[k], g) < 6e-2, k
        assert relerr(Gf[k], Gu[k]) < 3e-2, k
    assert relerr(df.float().cpu().reshape(12, 14, 14, 512).permute(0, 3, 1, 2), da_ref) < 6e-2


def test_fused_vocab_ce_label_smoothing_zero_and_full_length():
    cfg = dict(Bi=8, ncap=2, hw=(7, 7), D=512, A=128, E=256, H=512, V=1000, T=6, ragged=False)      # V not a multiple of 128
    W, ann, caps, lens = synth(**cfg, seed=14)
    ref = O.train_loss(W, ann, caps, lens, 0.0, 1.0)
    bf, Gf, _ = _fwd_bwd(W, ann, caps, lens, torch.bfloat16, True, fuse_ce=True, ls=0.0)
    assert abs(float(bf.t["out"][0]) - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))
    assert all(bool(torch.isfinite(v).all()) for v in Gf.values())


def test_bad_token_ids_are_flagged():
    cfg = dict(Bi=4, ncap=1, hw=(3, 3), D=64, A=32, E=32, H=64, V=128, T=5, ragged=False)
    W, ann, caps, lens = synth(**cfg, seed=15)
    buf, _, _ = _fwd_bwd(W, ann, caps, lens, torch.float32, False, fuse_ce=False)
    assert float(buf.t["out"][6]) == 0.0
    caps[1, 0, 2] = 1000                         # outside the vocabulary: nn.Embedding would raise
    buf, G, _ = _fwd_bwd(W, ann, caps, lens, torch.float32, False, fuse_ce=False)
    assert float(buf.t["out"][6]) == 1.0
    assert all(bool(torch.isfinite(v).all()) for v in G.values())
