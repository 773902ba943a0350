"""Further evaluators the reference's INIs name (reference: neuralmonkey/evaluators/{chrf,edit_distance,
wer,ter,average,mse}.py).  Each is a callable `(hypotheses, references) -> float` with a `name` and a
`compare_scores` (for error rates the smaller score is the better one).

TER needs the third-party `pyter` exactly as the reference does; without it the evaluator says so when
it is called rather than returning a different number."""
from collections import Counter
from difflib import SequenceMatcher
from typing import Callable, List, Sequence


def _check(hypotheses: Sequence, references: Sequence) -> None:
    if len(hypotheses) != len(references):
        raise ValueError("Hypothesis and reference lists do not have the same length: {} vs {}."
                         .format(len(hypotheses), len(references)))
    if not hypotheses:
        raise ValueError("No hyp/ref pair to evaluate.")


class _Metric:
    """Mean of a per-sentence score over the batch; `higher_is_better` fixes `compare_scores`."""
    higher_is_better = True

    def __init__(self, name: str) -> None:
        self.name = name

    def score_instance(self, hypothesis, reference) -> float:
        raise NotImplementedError

    def score_batch(self, hypotheses: List, references: List) -> float:
        _check(hypotheses, references)
        scores = [self.score_instance(h, r) for h, r in zip(hypotheses, references)]
        return float(sum(scores) / len(scores))

    def __call__(self, hypotheses: List, references: List) -> float:
        return self.score_batch(hypotheses, references)

    def compare_scores(self, score1: float, score2: float) -> int:
        first, second = (score1, score2) if self.higher_is_better else (score2, score1)
        return (first > second) - (first < second)


class ChrFEvaluator(_Metric):
    """chrF (Popovic 2015): F_beta of the character n-gram precision and recall, each averaged over the
    orders 1..n; an order without n-grams counts as 1 (evaluators/chrf.py:44-90).  Characters are those
    of the space-joined sentence, spaces included, minus `ignored_symbols`."""

    def __init__(self, n: int = 6, beta: float = 1.0, ignored_symbols: List[str] = None,
                 name: str = None) -> None:
        _Metric.__init__(self, name if name is not None else "ChrF-{}".format(beta))
        self.n, self.beta_2 = n, beta ** 2
        self.ignored = list(ignored_symbols) if ignored_symbols is not None else []

    def _orders(self, sentence: List[str]):
        chars = [c for c in " ".join(sentence) if c not in self.ignored]
        return chars, [Counter("".join(chars[i:i + m]) for i in range(len(chars) - m + 1))
                       for m in range(1, self.n + 1)]

    @staticmethod
    def _overlap(counted, other) -> float:
        ratios = []
        for own, theirs in zip(counted, other):
            total = sum(own.values())
            matched = sum(min(c, theirs[g]) for g, c in own.items() if g in theirs)
            ratios.append(matched / total if total else 1.0)
        return sum(ratios) / len(ratios)

    def chr_p(self, hyp_ngrams, ref_ngrams) -> float:
        """Character n-gram precision from per-order count dictionaries (evaluators/chrf.py:77-89)."""
        return self._overlap(hyp_ngrams, ref_ngrams)

    def chr_r(self, hyp_ngrams, ref_ngrams) -> float:
        """Character n-gram recall (evaluators/chrf.py:63-75)."""
        return self._overlap(ref_ngrams, hyp_ngrams)

    def score_instance(self, hypothesis: List[str], reference: List[str]) -> float:
        hyp_chars, hyp = self._orders(hypothesis)
        ref_chars, ref = self._orders(reference)
        if not hyp_chars or not ref_chars:
            return 1.0 if hyp_chars == ref_chars else 0.0
        precision, recall = self._overlap(hyp, ref), self._overlap(ref, hyp)
        if precision == 0.0 and recall == 0.0:
            return 0.0
        return (1 + self.beta_2) * precision * recall / (self.beta_2 * precision + recall)


class EditDistanceEvaluator(_Metric):
    """1 - mean difflib ratio of the space-joined sentences (evaluators/edit_distance.py)."""
    higher_is_better = False

    def score_instance(self, hypothesis: List[str], reference: List[str]) -> float:
        return SequenceMatcher(None, " ".join(hypothesis), " ".join(reference)).ratio()

    def score_batch(self, hypotheses, references) -> float:
        return 1 - _Metric.score_batch(self, hypotheses, references)


def _levenshtein(a: Sequence, b: Sequence) -> int:
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


class WEREvaluator(_Metric):
    """Word error rate: summed word-level edit distances over the summed reference lengths
    (evaluators/wer.py; the reference takes the distance from `pyter.edit_distance`, plain Levenshtein)."""
    higher_is_better = False

    def score_instance(self, hypothesis: List[str], reference: List[str]) -> float:
        if reference and hypothesis:
            return float(_levenshtein(hypothesis, reference))
        return 0.0 if not reference and not hypothesis else float(len(reference))

    def score_batch(self, hypotheses, references) -> float:
        _check(hypotheses, references)
        return (sum(self.score_instance(h, r) for h, r in zip(hypotheses, references))
                / sum(len(r) for r in references))


class TEREvaluator(_Metric):
    """Translation edit rate through `pyter.ter`, as evaluators/ter.py computes it."""
    higher_is_better = False

    def score_instance(self, hypothesis: List[str], reference: List[str]) -> float:
        if reference and hypothesis:
            try:
                import pyter
            except ImportError as exc:
                raise ImportError("The TER evaluator needs the 'pyter' package, as in Neural Monkey "
                                  "(requirements.txt); it is not installed.") from exc
            return pyter.ter(hypothesis, reference)
        return 0.0 if not reference and not hypothesis else 1.0


class AverageEvaluator(_Metric):
    """The mean of a runner's numeric outputs (evaluators/average.py)."""

    def score_instance(self, hypothesis: float, reference: float) -> float:
        return hypothesis


class PerplexityEvaluator(_Metric):
    """2 ** (sum of a runner's per-position cross-entropies / number of non-zero ones): masked positions carry a
    cross-entropy of exactly 0 and do not count (evaluators/perplexity.py; tests/language-model.ini feeds it the
    XentRunner's series).  NaN for a batch without any counted position.  The reference leaves `compare_scores`
    at the base class's "higher is better"; the INI sets `minimize_metric=True` instead."""

    def score_batch(self, hypotheses, references) -> float:
        _check(hypotheses, references)
        total = sum(float(x) for row in hypotheses for x in row)
        counted = sum(1 for row in hypotheses for x in row if x != 0.0)
        if counted == 0:
            return float("nan")
        return float(2 ** (total / counted))


class MeanSquaredErrorEvaluator(_Metric):
    """Mean of the element-wise squared errors of the whole batch (evaluators/mse.py:7-22)."""
    higher_is_better = False

    def score_batch(self, hypotheses, references) -> float:
        _check(hypotheses, references)
        errors = [(h - r) ** 2 for hyp, ref in zip(hypotheses, references) for h, r in zip(hyp, ref)]
        return float(sum(errors) / len(errors)) if errors else 0.0


class PairwiseMeanSquaredErrorEvaluator(_Metric):
    """Mean over the batch of each pair's own mean squared error (evaluators/mse.py:25-42)."""
    higher_is_better = False

    def score_instance(self, hypothesis: List[float], reference: List[float]) -> float:
        errors = [(h - r) ** 2 for h, r in zip(hypothesis, reference)]
        return float(sum(errors) / len(errors))


Metric = Callable[[List, List], float]
