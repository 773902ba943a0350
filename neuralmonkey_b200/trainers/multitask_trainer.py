"""Round-robin multi-task trainer (reference: neuralmonkey/trainers/multitask_trainer.py:12-47)."""
from typing import List

from neuralmonkey_b200.model.feedable import Feedable
from neuralmonkey_b200.runners.base_runner import GraphExecutor
from neuralmonkey_b200.trainers.generic_trainer import GenericTrainer


class MultitaskTrainer(GraphExecutor, Feedable):
    """Wraps several trainers and executes them one per batch, cyclically."""

    def __init__(self, trainers: List[GenericTrainer]) -> None:
        GraphExecutor.__init__(self, set(trainers))
        Feedable.__init__(self)
        self.trainers = trainers
        self.trainer_idx = 0

    @property
    def var_list(self):
        return list(set.union(*[set(t.var_list) for t in self.trainers])) if self.trainers else []

    @property
    def objectives(self):
        return [obj for t in self.trainers for obj in t.objectives]

    def get_executable(self, compute_losses: bool = True, summaries: bool = True,
                       num_sessions: int = 1):
        focused = self.trainers[self.trainer_idx]
        self.trainer_idx = (self.trainer_idx + 1) % len(self.trainers)
        return focused.get_executable(compute_losses, summaries, num_sessions)
