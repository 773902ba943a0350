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

        Each batch contributes the gradient of ITS token-mean loss (the reference's buffers sum the
        gradients of `train_loss` and divide by the number of batches, delayed_update_trainer.py:115-204).
        Data parallel: the mean of a batch is sum(xent over all ranks) / sum(mask over all ranks), so the
        two scalars of every micro-batch are exchanged right away (one 2-float all-reduce) and the
        gradients - of the un-normalised sum divided by the GLOBAL count - are exchanged once, on the update
        step; N ranks then reproduce the single-GPU update exactly."""
        import torch
        from neuralmonkey_b200 import distributed
        arena = runtime.arena()
        if arena.params.is_cuda:
            runtime.advance_dropout()       # new dropout masks for every micro-batch
        if self._accumulated == 0:
            arena.zero_grad()
        world = distributed.world_size()
        exact = None
        if len(self.objectives) == 1 and self.objectives[0].gradients is None:
            exact = self.objectives[0].loss_sum_and_count
        if exact is not None:
            loss_sum, count = exact
            stats = torch.stack([loss_sum.detach().reshape(()), count.detach().reshape(()).to(loss_sum.dtype)])
            distributed.all_reduce_sum(stats)
            w = self.objectives[0].weight
            term = loss_sum / stats[1]
            (term if w is None else term * w).backward()
            losses = [stats[0] / stats[1]]
            rank_scale = 1.0                # every rank already divided by the global count
        else:
            total = None
            for obj in self.objectives:
                if obj.gradients is not None:
                    raise NotImplementedError("objectives with explicit gradients (RL) are out of scope")
                w = 1.0 if obj.weight is None else obj.weight
                term = obj.loss * w
                total = term if total is None else total + term
            total.backward()
            losses = [obj.loss.detach() for obj in self.objectives]
            rank_scale = 1.0 / world        # mean of the per-rank means
        # gradients autograd delivered to plain torch expressions (e.g. an indexed embedding row) live in
        # `.grad` of the parameter views: fold them into the flat buffer after EVERY backward
        arena.fold_autograd_grads()
        self._accumulated += 1
        if self._accumulated == self.batches_per_update:
            distributed.all_reduce_sum(arena.allreduce_view)
            self.apply_gradients(rank_scale / self.batches_per_update, None)
            self._accumulated = 0
        return {"losses": losses, "l1l2": self._l1l2}
