"""Module path of the reference (neuralmonkey/evaluators/sacrebleu.py).  The `sacrebleu` package is not a
dependency here: the evaluator is the corpus BLEU of `evaluators/bleu.py` on the already tokenised series."""
from neuralmonkey_b200.evaluators import SacreBLEU  # noqa: F401
from neuralmonkey_b200.evaluators.bleu import BLEUEvaluator as SacreBLEUEvaluator  # noqa: F401
