"""Corpus BLEU (behaviour of neuralmonkey/evaluators/bleu.py: modified n-gram precision with
clipping, brevity penalty over the corpus, add-one-free, result in percent)."""
import math
from collections import Counter
from typing import List


class BLEUEvaluator:
    def __init__(self, n: int = 4, deduplicate: bool = False, name: str = None) -> None:
        self.n = n
        self.deduplicate = deduplicate
        self.name = name if name is not None else ("BLEU-{}".format(n) + ("-dedup" if deduplicate else ""))

    @staticmethod
    def ngrams(sentence: List[str], n: int) -> Counter:
        return Counter(tuple(sentence[i:i + n]) for i in range(len(sentence) - n + 1))

    @staticmethod
    def deduplicate_sentences(sentences: List[List[str]]) -> List[List[str]]:
        out = []
        for sent in sentences:
            dedup = []
            for tok in sent:
                if not dedup or dedup[-1] != tok:
                    dedup.append(tok)
            out.append(dedup)
        return out

    def __call__(self, decoded: List[List[str]], references: List[List[str]]) -> float:
        decoded, references = list(decoded), list(references)
        if self.deduplicate:
            decoded = self.deduplicate_sentences(decoded)
        log_prec = 0.0
        for n in range(1, self.n + 1):
            matched = total = 0
            for hyp, ref in zip(decoded, references):
                hyp_ng, ref_ng = self.ngrams(hyp, n), self.ngrams(ref, n)
                total += max(len(hyp) - n + 1, 0)
                matched += sum(min(c, ref_ng[g]) for g, c in hyp_ng.items())
            if matched == 0 or total == 0:
                return 0.0
            log_prec += math.log(matched / total) / self.n
        hyp_len = sum(len(h) for h in decoded)
        ref_len = sum(len(r) for r in references)
        if hyp_len == 0:
            return 0.0
        bp = 1.0 if hyp_len >= ref_len else math.exp(1.0 - ref_len / hyp_len)
        return 100.0 * bp * math.exp(log_prec)

    @staticmethod
    def compare_scores(score1: float, score2: float) -> int:
        return (score1 > score2) - (score1 < score2)


# pylint: disable=invalid-name
BLEU1 = BLEUEvaluator(n=1)
BLEU2 = BLEUEvaluator(n=2)
BLEU4 = BLEUEvaluator(n=4)
BLEU = BLEU4
BLEU4_dedup = BLEUEvaluator(n=4, deduplicate=True)
