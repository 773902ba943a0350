"""Module path of the reference (neuralmonkey/evaluators/wer.py) for INIs that name it; the classes live in
`evaluators/metrics.py`."""
from neuralmonkey_b200.evaluators.metrics import WEREvaluator  # noqa: F401

# pylint: disable=invalid-name
WER = WEREvaluator("WER")
