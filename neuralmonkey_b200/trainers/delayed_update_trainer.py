"""Gradient accumulation (reference: neuralmonkey/trainers/delayed_update_trainer.py:102-204):
gradients of `batches_per_update` consecutive batches are summed in the flat gradient arena,
then averaged, regularised, clipped and applied in one K13 launch."""
from typing import List

from neuralmonkey_b200 import runtime
from neuralmonkey_b200.trainers.generic_trainer import GenericTrainer
from neuralmonkey_b200.trainers.objective import Objective


class DelayedUpdateTrainer(GenericTrainer):
    # pylint: disable=too-many-arguments
    def __init__(self, batches_per_update: int, objectives: List[Objective], l1_weight: float = 0.0,
                 l2_weight: float = 0.0, clip_norm: float = None, optimizer=None,
                 var_scopes: List[str] = None, var_collection: str = None) -> None:
        GenericTrainer.__init__(self, objectives, l1_weight, l2_weight, clip_norm, optimizer,
                                var_scopes, var_collection)
        if batches_per_update < 1:
            raise ValueError("batches_per_update must be a positive integer")
        self.batches_per_update = batches_per_update
        self._accumulated = 0

    def train_step(self, apply_update: bool = True, grad_scale: float = 1.0, zero_grad: bool = True):
        """Accumulate; every `batches_per_update`-th call also applies the averaged update.
        (The per-batch token means are averaged, as the reference's buffers do.)"""
        arena = runtime.arena()
        first = self._accumulated == 0
        if first:
            arena.zero_grad()
        total = None
        for obj in self.objectives:
            w = 1.0 if obj.weight is None else obj.weight
            term = obj.loss * w
            total = term if total is None else total + term
        total.backward()
        losses = [obj.loss.detach() for obj in self.objectives]
        self._accumulated += 1
        if self._accumulated == self.batches_per_update:
            from neuralmonkey_b200 import distributed
            world = distributed.world_size()
            distributed.all_reduce_sum(arena.allreduce_view)
            self.apply_gradients(1.0 / (self.batches_per_update * world), None)
            self._accumulated = 0
        return {"losses": losses, "l1l2": self._l1l2}
