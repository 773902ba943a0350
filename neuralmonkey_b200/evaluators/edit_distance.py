"""Module path of the reference (neuralmonkey/evaluators/edit_distance.py) for INIs that name it; the classes live in
`evaluators/metrics.py`."""
from neuralmonkey_b200.evaluators.metrics import EditDistanceEvaluator  # noqa: F401

# pylint: disable=invalid-name
EditDistance = EditDistanceEvaluator("Edit distance")
