"""Sentence BLEU with NLTK's `SmoothingFunction().method2`, the one call the reference makes into
nltk (run_model.py:22,171,364: dev-time model selection and the per-commit BLEU it prints).  nltk
is not installable here, so the published algorithm is restated:

  p_n  = clipped n-gram matches / candidate n-grams, n = 1..4, uniform weights
  method2 (Lin & Och 2004): add 1 to numerator and denominator of p_n for n >= 2
  BLEU = BP * exp(sum_n 0.25 * log p_n),  BP = 1 if c > r else exp(1 - r/c)
  0 if there is no unigram match (nltk returns 0 before smoothing), 0 for an empty candidate.
"""
import math
from collections import Counter


def _ngrams(tokens, n):
    return Counter(tuple(tokens[i:i + n]) for i in range(len(tokens) - n + 1))


def sentence_bleu_method2(references, hypothesis, max_n=4):
    """references: list of token lists; hypothesis: token list."""
    c = len(hypothesis)
    if c == 0:
        return 0.0
    nums, dens = [], []
    for n in range(1, max_n + 1):
        hyp = _ngrams(hypothesis, n)
        best = Counter()
        for ref in references:
            for g, k in _ngrams(ref, n).items():
                best[g] = max(best[g], k)
        nums.append(sum(min(k, best[g]) for g, k in hyp.items()))
        dens.append(max(1, sum(hyp.values())))        # nltk: denominator floor 1 when the order is absent
    if nums[0] == 0:
        return 0.0
    r = min((abs(len(ref) - c), len(ref)) for ref in references)[1]
    bp = 1.0 if c > r else math.exp(1 - r / c)
    logs = 0.0
    for i in range(max_n):
        num, den = (nums[i], dens[i]) if i == 0 else (nums[i] + 1, dens[i] + 1)
        logs += 0.25 * math.log(num / den)
    return bp * math.exp(logs)
