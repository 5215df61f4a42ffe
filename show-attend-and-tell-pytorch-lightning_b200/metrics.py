"""Corpus BLEU / GLEU for SAT.score_captions (model.py:646-682) without nltk (not installed in the B200 image).

Restates the two third-party functions the reference imports at model.py:1-2, nltk 3.6.2 (requirements.txt:6):
  nltk.translate.bleu_score.corpus_bleu(list_of_references, hypotheses, weights)   -- Papineni et al. 2002: clipped n-gram
      precisions pooled over the corpus, closest-reference-length brevity penalty, no smoothing (a zero n-gram count is
      replaced by sys.float_info.min, as nltk's SmoothingFunction.method0 does);
  nltk.translate.gleu_score.corpus_gleu(list_of_references, hypotheses, min_len=1, max_len=4)  -- Wu et al. 2016: per
      hypothesis the reference with the best matching / max(|hyp n-grams|, |ref n-grams|) ratio, pooled over the corpus.
Sequences are lists of hashable tokens (word ids here).  Host-side integer work; known-answer tests: tests/test_metrics.py
(the values nltk's own docstrings give for their examples).
"""
import math
import sys
from collections import Counter
from fractions import Fraction


def _ngrams(seq, n):
    return [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]


def _everygrams(seq, min_len, max_len):
    out = []
    for n in range(min_len, max_len + 1):
        out.extend(_ngrams(seq, n))
    return out


def modified_precision(references, hypothesis, n):
    """(numerator, denominator) of the clipped n-gram precision of one hypothesis (nltk: modified_precision)."""
    counts = Counter(_ngrams(hypothesis, n)) if len(hypothesis) >= n else Counter()
    max_counts = {}
    for ref in references:
        ref_counts = Counter(_ngrams(ref, n)) if len(ref) >= n else Counter()
        for ng in counts:
            max_counts[ng] = max(max_counts.get(ng, 0), ref_counts[ng])
    clipped = {ng: min(c, max_counts[ng]) for ng, c in counts.items()}
    return sum(clipped.values()), max(1, sum(counts.values()))


def closest_ref_length(references, hyp_len):
    return min((len(r) for r in references), key=lambda rl: (abs(rl - hyp_len), rl))


def brevity_penalty(ref_len, hyp_len):
    if hyp_len > ref_len:
        return 1.0
    if hyp_len == 0:
        return 0.0
    return math.exp(1 - ref_len / hyp_len)


def corpus_bleu(list_of_references, hypotheses, weights=(0.25, 0.25, 0.25, 0.25)):
    num, den = Counter(), Counter()
    hyp_lengths = ref_lengths = 0
    assert len(list_of_references) == len(hypotheses), "The number of hypotheses and their reference(s) should be the same"
    for references, hypothesis in zip(list_of_references, hypotheses):
        for i in range(1, len(weights) + 1):
            a, b = modified_precision(references, hypothesis, i)
            num[i] += a
            den[i] += b
        hyp_lengths += len(hypothesis)
        ref_lengths += closest_ref_length(references, len(hypothesis))
    bp = brevity_penalty(ref_lengths, hyp_lengths)
    if num[1] == 0:
        return 0
    s = []
    for i, w in enumerate(weights, start=1):
        p = Fraction(num[i], den[i]) if num[i] != 0 else sys.float_info.min        # method0: no smoothing
        s.append(w * math.log(p))
    return bp * math.exp(math.fsum(s))


def sentence_bleu(references, hypothesis, weights=(0.25, 0.25, 0.25, 0.25)):
    return corpus_bleu([references], [hypothesis], weights)


def corpus_gleu(list_of_references, hypotheses, min_len=1, max_len=4):
    assert len(list_of_references) == len(hypotheses), "The number of hypotheses and their reference(s) should be the same"
    corpus_n_match = corpus_n_all = 0
    for references, hypothesis in zip(list_of_references, hypotheses):
        hyp_ngrams = Counter(_everygrams(hypothesis, min_len, max_len))
        tpfp = sum(hyp_ngrams.values())
        hyp_counts = []
        for reference in references:
            ref_ngrams = Counter(_everygrams(reference, min_len, max_len))
            tpfn = sum(ref_ngrams.values())
            tp = sum((ref_ngrams & hyp_ngrams).values())
            n_all = max(tpfp, tpfn)
            if n_all > 0:
                hyp_counts.append((tp, n_all))
        if hyp_counts:
            n_match, n_all = max(hyp_counts, key=lambda hc: hc[0] / hc[1])
            corpus_n_match += n_match
            corpus_n_all += n_all
    return 0.0 if corpus_n_all == 0 else corpus_n_match / corpus_n_all


def sentence_gleu(references, hypothesis, min_len=1, max_len=4):
    return corpus_gleu([references], [hypothesis], min_len, max_len)
