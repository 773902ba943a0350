"""SacreBLEU evaluator (reference: neuralmonkey/evaluators/sacrebleu.py, a wrapper over
`sacrebleu.corpus_bleu`).  The `sacrebleu` package is not available here, so the part of it the wrapper
uses is restated: corpus BLEU over 1..4-grams with clipped counts, brevity penalty on the reference length,
the smoothing methods "exp" (NIST mteval: a precision with zero matches becomes 1 / (2^k * total), k counting
such orders), "floor" and "none", optional lowercasing, the tokenizers "none" (the wrapper's default: the
series is already tokenised, tokens are joined and split on spaces) and "13a" (mteval-v13a), and
`use_effective_order`.  The score is on sacrebleu's 0-100 scale."""
import math
import re
from collections import Counter
from typing import List

SMOOTH_VARIANTS = ["exp", "floor", "none"]
TOKENIZERS = ["none", "13a"]
NGRAM_ORDER = 4


def tokenize_13a(line: str) -> str:
    """mteval-v13a tokenisation as sacrebleu applies it."""
    norm = line.replace("<skipped>", "").replace("-\n", "").replace("\n", " ")
    norm = norm.replace("&quot;", '"').replace("&amp;", "&").replace("&lt;", "<").replace("&gt;", ">")
    norm = " {} ".format(norm)
    norm = re.sub(r"([\{-\~\[-\` -\&\(-\+\:-\@\/])", r" \1 ", norm)
    norm = re.sub(r"([^0-9])([\.,])", r"\1 \2 ", norm)
    norm = re.sub(r"([\.,])([^0-9])", r" \1 \2", norm)
    norm = re.sub(r"([0-9])(-)", r"\1 \2 ", norm)
    return " ".join(norm.split())


def _ngrams(tokens: List[str]) -> Counter:
    counts = Counter()  # type: Counter
    for n in range(1, NGRAM_ORDER + 1):
        for i in range(len(tokens) - n + 1):
            counts[tuple(tokens[i:i + n])] += 1
    return counts


class SacreBLEUEvaluator:
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, smooth_method: str = "exp", smooth_value: float = 0.0, force: bool = False,
                 lowercase: bool = False, tokenize: str = "none", use_effective_order: bool = False) -> None:
        if tokenize not in TOKENIZERS:
            raise ValueError("Unknown tokenizer '{}'. You must use one of sacrebleu's tokenizers: {}"
                             .format(tokenize, str(TOKENIZERS)))
        if smooth_method not in SMOOTH_VARIANTS:
            raise ValueError("Unknown smoothing '{}'. You must use one of sacrebleu's smoothing methods: {}"
                             .format(smooth_method, str(SMOOTH_VARIANTS)))
        self.name = name
        self.smooth_method = smooth_method
        self.smooth_value = smooth_value
        self.force = force
        self.lowercase = lowercase
        self.tokenize = tokenize
        self.use_effective_order = use_effective_order

    def _prepare(self, sentence: List[str]) -> List[str]:
        line = " ".join(sentence)
        if self.lowercase:
            line = line.lower()
        if self.tokenize == "13a":
            line = tokenize_13a(line)
        return line.split()

    def score_batch(self, hypotheses: List[List[str]], references: List[List[str]]) -> float:
        if len(hypotheses) != len(references):
            raise ValueError("Hypothesis and reference lists do not have the same length: {} vs {}."
                             .format(len(hypotheses), len(references)))
        correct, total = [0] * NGRAM_ORDER, [0] * NGRAM_ORDER
        sys_len = ref_len = 0
        for hyp, ref in zip(hypotheses, references):
            hyp_t, ref_t = self._prepare(hyp), self._prepare(ref)
            sys_len += len(hyp_t)
            ref_len += len(ref_t)
            ref_counts = _ngrams(ref_t)
            for gram, count in _ngrams(hyp_t).items():
                total[len(gram) - 1] += count
                correct[len(gram) - 1] += min(count, ref_counts.get(gram, 0))
        precisions = [0.0] * NGRAM_ORDER
        smooth_mteval, effective_order = 1.0, NGRAM_ORDER
        for n in range(NGRAM_ORDER):
            if total[n] == 0:
                break
            if self.use_effective_order:
                effective_order = n + 1
            if correct[n] == 0:
                if self.smooth_method == "exp":
                    smooth_mteval *= 2
                    precisions[n] = 100.0 / (smooth_mteval * total[n])
                elif self.smooth_method == "floor":
                    precisions[n] = 100.0 * self.smooth_value / total[n]
            else:
                precisions[n] = 100.0 * correct[n] / total[n]
        if sys_len < ref_len:
            brevity = math.exp(1 - ref_len / sys_len) if sys_len > 0 else 0.0
        else:
            brevity = 1.0
        logs = [math.log(p) if p > 0 else -9999999999.0 for p in precisions[:effective_order]]
        return brevity * math.exp(sum(logs) / effective_order)

    def __call__(self, decoded: List[List[str]], references: List[List[str]]) -> float:
        return self.score_batch(decoded, references)

    @staticmethod
    def compare_scores(score1: float, score2: float) -> int:
        return (score1 > score2) - (score1 < score2)


# pylint: disable=invalid-name
SacreBLEU = SacreBLEUEvaluator("BLEU")
