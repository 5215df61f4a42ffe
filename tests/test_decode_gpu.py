"""GPU parity of batched greedy / beam decoding: bit-exact token ids against the reference goldens and the
CPU oracle (fp32 mode), scores / perplexities / alphas within 1e-5 relative."""
import pytest
import torch

from conftest import load_golden
from oracle import sat_oracle as O
from test_train_forward_gpu import relerr

pytestmark = pytest.mark.gpu

VOC = lambda V: dict(PAD=0, UNK=V - 3, START=V - 2, END=V - 1)


def cuda_caption(W, ann, k, max_len, temperature=1.0, rescore=None, reward=0.5, return_all=False, dtype=torch.float32,
                 use_tc=False):
    from sat_b200 import decode, decoder
    V = W["embedding.weight"].shape[0]
    dw = decode.DecodeWeights(W, dtype, torch.device("cuda"), dtype == torch.float32, use_tc)
    bld = decoder.annotations_as_bld(ann.cuda(), dtype)
    t = decode.decode_annotations(dw, bld, k, max_len, temperature, rescore, reward, VOC(V))
    return decode.assemble(t, tuple(ann.shape[2:]), return_all=return_all)


@pytest.mark.parametrize("k", [1, 3, 5])
@pytest.mark.parametrize("rescore", [None, "LN", "WR", "BAR"])
@pytest.mark.parametrize("return_all", [False, True])
def test_decode_vs_reference_golden(k, rescore, return_all):
    z, W, _ = load_golden("decode_small")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    caps, scores, alphas, ppl = cuda_caption(W, ann, k, max_len, 1.0, rescore, 0.5, return_all)
    tag = "k%d_%s_%s" % (k, rescore, "all" if return_all else "best")
    for i in range(ann.shape[0]):
        cc, ss, aa, pp = (caps[i], scores[i], alphas[i], ppl[i]) if return_all else ([caps[i]], [scores[i]], [alphas[i]], [ppl[i]])
        assert len(cc) == int(z["%s/n%d/count" % (tag, i)])
        for j in range(len(cc)):
            assert cc[j] == z["%s/n%d/h%d/tokens" % (tag, i, j)].tolist()
            ref_s = float(z["%s/n%d/h%d/score" % (tag, i, j)])
            assert abs(ss[j] - ref_s) < 1e-5 * max(1.0, abs(ref_s))
            ref_p = float(z["%s/n%d/h%d/ppl" % (tag, i, j)])
            assert abs(pp[j] - ref_p) < 1e-4 * max(1.0, abs(ref_p))
            assert tuple(aa[j].shape) == tuple(z["%s/n%d/h%d/alphas" % (tag, i, j)].shape)
            assert relerr(aa[j], z["%s/n%d/h%d/alphas" % (tag, i, j)]) < 1e-5


def test_decode_temperature_vs_reference_golden():
    z, W, _ = load_golden("decode_small")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    caps, scores, _, _ = cuda_caption(W, ann, 3, max_len, 0.7, "LN")
    for i in range(ann.shape[0]):
        assert caps[i] == z["k3_LN_T0.7/n%d/tokens" % i].tolist()
        assert abs(scores[i] - float(z["k3_LN_T0.7/n%d/score" % i])) < 1e-4


@pytest.mark.parametrize("k", [1, 5])
def test_decode_c1_dims_vs_oracle(k):
    """BASELINE decoder dims (L=196, D=512, A=128, E=256, H=512, V=6400), sharpened weights so that <END> and the
    shrinking beam are exercised; reports the minimum top-2 margin of the greedy run."""
    D, A, E, H, V = 512, 128, 256, 512, 6400
    W = O.random_weights(D, A, E, H, V, seed=11, sharpen=True)
    g = torch.Generator().manual_seed(12)
    ann = torch.randn(6, D, 14, 14, generator=g)
    ref = O.caption(W, ann, VOC(V), beamk=k, max_gen_length=30, rescore_method="LN", return_all=True)
    got = cuda_caption(W, ann, k, 30, 1.0, "LN", 0.5, True)
    assert got[0] == ref[0]
    for a, b in zip(got[1], ref[1]):
        assert max(abs(x - y) for x, y in zip(a, b)) < 1e-4
    lens = sorted(len(c) for cc in ref[0] for c in cc)
    print("caption lengths", lens)


@pytest.mark.parametrize("use_tc", [False, True])
def test_greedy_bf16_runs_and_mostly_agrees(use_tc):
    D, A, E, H, V = 512, 128, 256, 512, 6400
    W = O.random_weights(D, A, E, H, V, seed=11, sharpen=True)
    g = torch.Generator().manual_seed(12)
    ann = torch.randn(6, D, 14, 14, generator=g)
    ref = O.caption(W, ann, VOC(V), beamk=1, max_gen_length=30)
    got = cuda_caption(W, ann, 1, 30, dtype=torch.bfloat16, use_tc=use_tc)
    agree = sum(1 for a, b in zip(got[0], ref[0]) if a[:3] == b[:3])
    assert agree >= 3          # bf16 token ids are not expected to be bit-exact ...
    from test_full_size_gpu import oracle_gaps
    gaps = oracle_gaps(W, ann, got[0], VOC(V))
    assert max(max(gp) for gp in gaps) < 0.1      # ... but every chosen word is a near-tie (nats) under the teacher-forced fp32 oracle


def test_decode_single_image_and_wide_beam():
    """one image, beam wider than typical (k=8) and k=1, tiny vocabulary: exercises shrinking beams down to zero."""
    D, A, E, H, V = 64, 32, 32, 64, 64
    W = O.random_weights(D, A, E, H, V, seed=21, sharpen=True)
    W["output.output.bias"][V - 1] = 2.0
    g = torch.Generator().manual_seed(22)
    ann = torch.randn(1, D, 3, 3, generator=g)
    for k in (1, 8):
        ref = O.caption(W, ann, VOC(V), beamk=k, max_gen_length=12, rescore_method="LN", return_all=True)
        got = cuda_caption(W, ann, k, 12, 1.0, "LN", 0.5, True)
        assert got[0] == ref[0]
        assert max(abs(x - y) for x, y in zip(got[1][0], ref[1][0])) < 1e-4
        assert len(got[0][0]) == k            # every beam ends up as exactly one finished hypothesis


def test_decode_temperature_list_cycles_per_step():
    z, W, _ = load_golden("decode_small")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    temps = [1.0, 0.6, 1.4]
    ref = O.caption(W, ann, VOC(V), beamk=3, max_gen_length=max_len, temperature=temps, rescore_method="WR")
    got = cuda_caption(W, ann, 3, max_len, temps, "WR")
    assert got[0] == ref[0]
    assert max(abs(x - y) for x, y in zip(got[1], ref[1])) < 1e-4


@pytest.mark.parametrize("n_tied", [20, 390])
def test_beam_candidates_with_tied_logits(n_tied):
    """Tied logits: candidates are taken in (value desc, vocabulary index asc) order.  20 ties stay inside the
    threshold kernel's candidate list, 390 overflow it and take the block-wide scan."""
    D, A, E, H, V, k = 64, 32, 32, 64, 1024, 3
    W = O.random_weights(D, A, E, H, V, seed=31)
    W["output.output.weight"].zero_()
    W["output.output.bias"].zero_()
    W["output.output.bias"][10:10 + n_tied] = 1.0
    g = torch.Generator().manual_seed(32)
    ann = torch.randn(2, D, 3, 3, generator=g)
    caps, scores, _, _ = cuda_caption(W, ann, k, 6, 1.0, None, 0.5, True)
    for i in range(2):
        assert len(caps[i]) == k
        for c in caps[i]:
            assert len(c) == 6 and set(c) <= {10, 11, 12}, c           # never <END>: flushed at the maximum length
        assert all(abs(s - scores[i][0]) < 1e-5 for s in scores[i])     # every hypothesis has the same score


def test_rows_that_die_mid_decode_produce_zero_attention():
    """A beam slot that stops being live (its hypothesis ended) is skipped by every kernel of the following steps: the
    attention kernel reads the liveness array only after its predecessors in the launch chain have completed (it is
    rewritten every step by beam_update_kernel), and writes exact zeros for dead slots."""
    from sat_b200 import decode, decoder
    z, W, _ = load_golden("decode_small")
    V, max_len = int(z["dims"][4]), int(z["dims"][5])
    ann = torch.from_numpy(z["ann"])
    k = 5
    dw = decode.DecodeWeights(W, torch.float32, torch.device("cuda"), True, False)
    t = decode.decode_annotations(dw, decoder.annotations_as_bld(ann.cuda(), torch.float32), k, max_len, 1.0, "LN", 0.5, VOC(V))
    torch.cuda.synchronize()
    alpha_all = t["alpha_all"].cpu()                       # [S+1, R, L]
    fin_len = t["fin_len"].cpu()
    fin_cnt = t["fin_count"].cpu()
    n_img = ann.shape[0]
    checked_dead = 0
    for n in range(n_img):
        lens_n = sorted(int(fin_len[n, i]) for i in range(int(fin_cnt[n])))
        for step in range(1, max_len + 1):
            # a hypothesis of length l (fin_len = step at which it ended) leaves the beam after step l; flushed ones end at max_len
            live = k - sum(1 for l in lens_n if l < step)
            for j in range(k):
                row = alpha_all[step, n * k + j]
                if j < live:
                    assert abs(float(row.sum()) - 1.0) < 1e-5, (n, step, j)
                else:
                    assert float(row.abs().max()) == 0.0, (n, step, j)
                    checked_dead += 1
    assert checked_dead > 0                                # the golden's beams do shrink


def test_early_out_stops_the_launch_loop_and_keeps_results():
    """model.py:419,436: the reference leaves an image's loop when its beams are used up.  Here the kernel that retires the last
    live image raises a host-visible flag and the launch loop stops queueing steps (it is the host that paces the loop for a
    few images).  Same captions / scores as the full-length run, fewer launches."""
    from sat_b200 import _lib, decode, decoder
    D, A, E, H, V = 64, 32, 32, 64, 64
    W = O.random_weights(D, A, E, H, V, seed=31, sharpen=True)
    W["output.output.bias"][V - 1] = 30.0                       # <END> wins as soon as it is admissible (step 1)
    g = torch.Generator().manual_seed(32)
    ann = torch.randn(2, D, 3, 3, generator=g)
    dw = decode.DecodeWeights(W, torch.float32, torch.device("cuda"), True, False)
    bld = decoder.annotations_as_bld(ann.cuda(), torch.float32)
    out, launches = {}, {}
    for early in (False, True):
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        t = decode.decode_annotations(dw, bld, 3, 60, 1.0, "LN", 0.5, VOC(V), early_out=early)
        launches[early] = _lib.launch_count() - n0
        out[early] = decode.assemble(t, (3, 3), return_all=True)
    assert out[True][0] == out[False][0] and out[True][1] == out[False][1]
    assert max(len(c) for img in out[True][0] for c in img) <= 3
    ref = O.caption(W, ann, VOC(V), beamk=3, max_gen_length=60, rescore_method="LN", return_all=True)
    assert out[True][0] == ref[0]
    assert launches[True] < launches[False], launches


def test_train_and_decode_alternating_shapes_share_the_attention_kernel():
    """regression: training with a large annotation map, decoding with a small one, training again -- the shared-memory opt-in
    of the attention kernel is one attribute for the whole library, whichever driver launched last"""
    from test_train_backward_gpu import run_cuda_fwd_bwd
    from test_train_forward_gpu import synth
    big = dict(Bi=4, ncap=1, hw=(14, 14), D=512, A=128, E=64, H=64, V=128, T=4, ragged=True)
    small = dict(Bi=4, ncap=1, hw=(12, 12), D=512, A=128, E=64, H=64, V=128, T=4, ragged=True)
    Wb, annb, capsb, lensb = synth(**big)
    Ws, anns, _, _ = synth(**small, sharpen=True)
    l1, _, _ = run_cuda_fwd_bwd(Wb, annb, capsb, lensb, 0.0, 1.0)
    cuda_caption(Ws, anns, 1, 6)
    l2, _, _ = run_cuda_fwd_bwd(Wb, annb, capsb, lensb, 0.0, 1.0)
    assert l1 == l2
