"""Evaluators referenced by the target configs (reference: neuralmonkey/evaluators/).
BLEU is implemented natively; SacreBLEU restates what the reference's wrapper uses of the `sacrebleu` package
(not installed here): evaluators/sacrebleu.py; ROUGE-L is the LCS F-score."""
from typing import List

from neuralmonkey_b200.evaluators.bleu import BLEU, BLEU1, BLEU2, BLEU4, BLEUEvaluator
from neuralmonkey_b200.evaluators.metrics import (AverageEvaluator, ChrFEvaluator, EditDistanceEvaluator,
                                                  MeanSquaredErrorEvaluator,
                                                  PairwiseMeanSquaredErrorEvaluator, PerplexityEvaluator,
                                                  TEREvaluator,
                                                  WEREvaluator)


class AccuracyEvaluator:
    """Token-level accuracy: aligned hypothesis / reference tokens of the whole batch are scored
    with `==` and averaged together (evaluators/evaluator.py:122-187, evaluators/accuracy.py);
    reference tokens equal to `mask_symbol` are left out."""

    def __init__(self, name: str = "Accuracy", mask_symbol=None) -> None:
        self.name = name
        self.mask_symbol = mask_symbol

    def __call__(self, decoded: List, references: List) -> float:
        hits, total = 0.0, 0
        for hyp, ref in zip(decoded, references):
            for h_tok, r_tok in zip(hyp, ref):
                if self.mask_symbol and r_tok == self.mask_symbol:
                    continue
                hits += float(h_tok == r_tok)
                total += 1
        return hits / total if total else 0.0


class AccuracySeqLevelEvaluator:
    """1.0 for an exactly matching sequence, 0.0 otherwise, averaged over the batch."""

    def __init__(self, name: str = "AccuracySeqLevel") -> None:
        self.name = name

    def __call__(self, decoded: List, references: List) -> float:
        pairs = list(zip(decoded, references))
        return sum(1.0 for hyp, ref in pairs if hyp == ref) / len(pairs) if pairs else 0.0


class RougeLEvaluator:
    """Sentence-level ROUGE-L F-score averaged over the corpus (evaluators/rouge.py)."""

    def __init__(self, name: str = "ROUGE-L", beta: float = 1.2) -> None:
        self.name = name
        self.beta = beta

    @staticmethod
    def _lcs(a: List[str], b: List[str]) -> int:
        prev = [0] * (len(b) + 1)
        for x in a:
            cur = [0]
            for j, y in enumerate(b):
                cur.append(prev[j] + 1 if x == y else max(prev[j + 1], cur[j]))
            prev = cur
        return prev[-1]

    def __call__(self, decoded: List[List[str]], references: List[List[str]]) -> float:
        scores = []
        for hyp, ref in zip(decoded, references):
            lcs = self._lcs(list(hyp), list(ref))
            if lcs == 0 or not hyp or not ref:
                scores.append(0.0)
                continue
            p, r = lcs / len(hyp), lcs / len(ref)
            scores.append((1 + self.beta ** 2) * p * r / (r + self.beta ** 2 * p))
        return sum(scores) / len(scores) if scores else 0.0


# pylint: disable=invalid-name
Accuracy = AccuracyEvaluator()
ROUGE_L = RougeLEvaluator()
# the reference names this instance "BLEU" (evaluators/sacrebleu.py:63): its results are logged as <series>/BLEU
from neuralmonkey_b200.evaluators.sacrebleu import SacreBLEU, SacreBLEUEvaluator  # noqa: E402
AccuracySeqLevel = AccuracySeqLevelEvaluator()
from neuralmonkey_b200.evaluators.chrf import ChrF3  # noqa: E402
from neuralmonkey_b200.evaluators.edit_distance import EditDistance  # noqa: E402
from neuralmonkey_b200.evaluators.mse import MSE, PairwiseMSE  # noqa: E402
from neuralmonkey_b200.evaluators.ter import TER  # noqa: E402
from neuralmonkey_b200.evaluators.wer import WER  # noqa: E402
