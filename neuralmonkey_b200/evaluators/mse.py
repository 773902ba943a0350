"""Module path of the reference (neuralmonkey/evaluators/mse.py) for INIs that name it; the classes live in
`evaluators/metrics.py`."""
from neuralmonkey_b200.evaluators.metrics import MeanSquaredErrorEvaluator, PairwiseMeanSquaredErrorEvaluator  # noqa: F401

# pylint: disable=invalid-name
MSE = MeanSquaredErrorEvaluator("MeanSquaredError")
PairwiseMSE = PairwiseMeanSquaredErrorEvaluator("PairwiseMeanSquaredError")
