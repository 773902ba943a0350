"""Cross-entropy trainer (reference: neuralmonkey/trainers/cross_entropy_trainer.py:21-53)."""
from typing import Any, List

from neuralmonkey_b200.logging import warn
from neuralmonkey_b200.trainers.generic_trainer import GenericTrainer
from neuralmonkey_b200.trainers.objective import CostObjective, Objective, ObjectiveWeight


def xent_objective(decoder, weight=None) -> Objective:
    warn("Using deprecated xent_objective function. Use the CostObjective class directly.")
    return CostObjective(decoder, weight)


class CrossEntropyTrainer(GenericTrainer):
    # pylint: disable=too-many-arguments
    def __init__(self, decoders: List[Any], decoder_weights: List[ObjectiveWeight] = None,
                 l1_weight: float = 0., l2_weight: float = 0., clip_norm: float = None,
                 optimizer=None, var_scopes: List[str] = None, var_collection: str = None,
                 use_cuda_graph: bool = False) -> None:
        if decoder_weights is None:
            decoder_weights = [None for _ in decoders]
        if len(decoder_weights) != len(decoders):
            raise ValueError("decoder_weights (length {}) do not match decoders (length {})"
                             .format(len(decoder_weights), len(decoders)))
        objectives = [CostObjective(dec, w) for dec, w in zip(decoders, decoder_weights)]
        GenericTrainer.__init__(self, objectives=objectives, l1_weight=l1_weight,
                                l2_weight=l2_weight, clip_norm=clip_norm, optimizer=optimizer,
                                var_scopes=var_scopes, var_collection=var_collection,
                                use_cuda_graph=use_cuda_graph)
