"""Module path of the reference (neuralmonkey/evaluators/ter.py) for INIs that name it; the classes live in
`evaluators/metrics.py`."""
from neuralmonkey_b200.evaluators.metrics import TEREvaluator  # noqa: F401

# pylint: disable=invalid-name
TER = TEREvaluator("TER")
