"""Feedable: receives one batch of data series (reference: neuralmonkey/model/feedable.py).

`feed_dict(dataset, train)` converts this part's series to device tensors, stores them on
the part and invalidates every per-batch `@tensor` cache.  `batch_size` and `train_mode`
are the two inputs every feedable gets (feedable.py:28-50).
"""
from typing import Any, Dict


class Feedable:
    def __init__(self) -> None:
        self.train_mode = False
        self.batch_size = 0
        self._inputs = {}  # type: Dict[str, Any]

    def reset_batch(self) -> None:
        self.__dict__["_batch_cache"] = {}

    def feed_dict(self, dataset, train: bool = True) -> Dict[str, Any]:
        self.reset_batch()
        self.train_mode = bool(train)
        self.batch_size = len(dataset)
        self._inputs = {}
        return self._inputs

    @property
    def input_types(self) -> Dict[str, Any]:
        return {}

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {}

    @property
    def dataset(self) -> Dict[str, Any]:
        return self._inputs

    # -- CUDA-graph support (trainers/generic_trainer.py) ------------------------------------------
    def static_inputs(self) -> Dict[str, Any]:
        """Device tensors this part received from the feed stage (its graph leaves): everything
        else it exposes is computed from them.  Default: none."""
        return {}

    def bind_static(self, tensors: Dict[str, Any]) -> None:
        """Drop every per-batch cache and read the leaves from `tensors` (same keys) instead."""
        self.reset_batch()

    def register_input(self, dataset: Dict[str, Any]) -> None:
        """Kept for API compatibility (experiment.py:152-174): inputs are fed per batch."""
