"""Module path of the reference (neuralmonkey/evaluators/accuracy.py)."""
from neuralmonkey_b200.evaluators import (Accuracy, AccuracyEvaluator, AccuracySeqLevel,  # noqa: F401
                                          AccuracySeqLevelEvaluator)
