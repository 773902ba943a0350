"""Corpus BLEU with the behaviour of neuralmonkey/evaluators/bleu.py, quirks included:

* n-grams are space-joined strings; several references per sentence may be packed into one token list
  with a separator token, their n-gram counts merged by maximum;
* the "modified precision" adds, for every DISTINCT hypothesis n-gram, its (merged) reference count -
  not min(hypothesis count, reference count) (:113-116);
* an order without any hypothesis n-gram has precision 1; an order with zero matches is smoothed as in
  mteval-v13a: smooth *= 2, precision = 1 / (smooth * number of hypothesis n-grams) (:196-204);
* brevity penalty exp(min(1 - r/c, 0)) with r the reference length closest to each hypothesis' length;
  an empty hypothesis corpus scores 0;
* the result is in percent."""
import math
from collections import Counter
from typing import List, Optional, Tuple


class BLEUEvaluator:
    def __init__(self, n: int = 4, deduplicate: bool = False, name: str = None,
                 multiple_references_separator: Optional[str] = None) -> None:
        self.n = n
        self.deduplicate = deduplicate
        self.multiple_references_separator = multiple_references_separator
        self.name = name if name is not None else ("BLEU-{}".format(n) + ("-dedup" if deduplicate else ""))

    # -- pieces -----------------------------------------------------------------------------------
    @staticmethod
    def ngram_counts(sentence: List[str], n: int, lowercase: bool = False, delimiter: str = " ") -> Counter:
        grams = (delimiter.join(sentence[i:i + n]) for i in range(len(sentence) - n + 1))
        return Counter(g.lower() if lowercase else g for g in grams)

    @staticmethod
    def merge_max_counters(counters: List[Counter]) -> Counter:
        merged = Counter()  # type: Counter
        for counter in counters:
            for key, value in counter.items():
                merged[key] = max(merged[key], value)
        return merged

    @staticmethod
    def modified_ngram_precision(hypotheses: List[List[str]], references_list: List[List[List[str]]], n: int,
                                 case_sensitive: bool = True) -> Tuple[float, int]:
        matched = generated = 0
        for hypothesis, references in zip(hypotheses, references_list):
            allowed = BLEUEvaluator.merge_max_counters(
                [BLEUEvaluator.ngram_counts(ref, n, not case_sensitive) for ref in references])
            produced = BLEUEvaluator.ngram_counts(hypothesis, n, not case_sensitive)
            matched += sum(allowed[gram] for gram in produced)
            generated += sum(produced.values())
        return (1, 0) if generated == 0 else (matched / generated, generated)

    @staticmethod
    def effective_reference_length(hypotheses: List[List[str]], references_list: List[List[List[str]]]) -> int:
        total = 0
        for hypothesis, references in zip(hypotheses, references_list):
            best_diff, best_length = math.inf, 0
            for reference in references:                 # the first of equally close references wins
                diff = abs(len(reference) - len(hypothesis))
                if diff < best_diff:
                    best_diff, best_length = diff, len(reference)
            total += best_length
        return total

    @staticmethod
    def bleu(hypotheses: List[List[str]], references: List[List[List[str]]], ngrams: int = 4,
             case_sensitive: bool = True) -> float:
        log_bleu, smooth = 0.0, 1.0
        for order in range(1, ngrams + 1):
            precision, generated = BLEUEvaluator.modified_ngram_precision(hypotheses, references, order,
                                                                          case_sensitive)
            if precision == 0:
                smooth *= 2
                precision = 1 / (smooth * generated)
            log_bleu += math.log(precision) / ngrams
        ref_length = BLEUEvaluator.effective_reference_length(hypotheses, references)
        hyp_length = sum(len(h) for h in hypotheses)
        if hyp_length == 0:
            return 0.0
        return math.exp(log_bleu + min(1 - ref_length / hyp_length, 0))

    @staticmethod
    def deduplicate_sentences(sentences: List[List[str]]) -> List[List[str]]:
        return [[tok for i, tok in enumerate(sent) if i == 0 or sent[i - 1] != tok] for sent in sentences]

    def _split_references(self, packed: List[str]) -> List[List[str]]:
        if self.multiple_references_separator is None:
            return [packed]
        references, current = [], []  # type: List[List[str]], List[str]
        for tok in packed:
            if tok == self.multiple_references_separator:
                references.append(current)
                current = []
            else:
                current.append(tok)
        return references + [current]

    # -- evaluator protocol -----------------------------------------------------------------------
    def score_batch(self, hypotheses: List[List[str]], references: List[List[str]]) -> float:
        hypotheses, references = list(hypotheses), list(references)
        if len(hypotheses) != len(references):
            raise ValueError("Hypothesis and reference lists do not have the same length: {} vs {}."
                             .format(len(hypotheses), len(references)))
        if not hypotheses:
            raise ValueError("No hyp/ref pair to evaluate.")
        listed = [self._split_references(list(r)) for r in references]
        if self.deduplicate:
            hypotheses = self.deduplicate_sentences(hypotheses)
        return 100 * self.bleu([list(h) for h in hypotheses], listed, self.n)

    def __call__(self, decoded: List[List[str]], references: List[List[str]]) -> float:
        return self.score_batch(decoded, references)

    @staticmethod
    def compare_scores(score1: float, score2: float) -> int:
        return (score1 > score2) - (score1 < score2)


# pylint: disable=invalid-name
BLEU1 = BLEUEvaluator(n=1)
BLEU2 = BLEUEvaluator(n=2)
BLEU4 = BLEUEvaluator(n=4)
BLEU = BLEUEvaluator()
BLEU4_dedup = BLEUEvaluator(n=4, deduplicate=True)
