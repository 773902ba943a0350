"""Module path of the reference (neuralmonkey/evaluators/rouge.py); ROUGE-L only (ROUGE-1 / ROUGE-2 come from
the third-party `rouge` package there)."""
from neuralmonkey_b200.evaluators import ROUGE_L, RougeLEvaluator  # noqa: F401
