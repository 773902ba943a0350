"""Module path of the reference (neuralmonkey/evaluators/chrf.py) for INIs that name it; the classes live in
`evaluators/metrics.py`."""
from neuralmonkey_b200.evaluators.metrics import ChrFEvaluator  # noqa: F401

# pylint: disable=invalid-name
ChrF3 = ChrFEvaluator(beta=3)


def _get_ngrams(tokens, n):
    """Character n-gram counts of orders 1..n as a list of dictionaries (evaluators/chrf.py:91-100)."""
    from collections import Counter
    return [dict(Counter("".join(tokens[i:i + m]) for i in range(len(tokens) - m + 1))) for m in range(1, n + 1)]
