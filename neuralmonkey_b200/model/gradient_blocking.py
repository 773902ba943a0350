"""Gradient-blocking views of encoder-like objects (reference: neuralmonkey/model/gradient_blocking.py;
tests/bpe.ini freezes its encoder this way in the second run of tests_run.sh): the same states with the
graph cut behind them, so nothing upstream receives a gradient through the view.  A view is not a model part -
it owns no variables and is fed nothing - but the object it wraps stays reachable through `dependencies`, so
its variables and feeds are still collected.  The values are per-batch tensors of the wrapped object; the
view re-reads them at every access (a `detach()` is a view of the same memory)."""
from typing import List

from neuralmonkey_b200.model.stateful import SpatialStateful, Stateful, TemporalStateful
from neuralmonkey_b200.typecheck import check_argument_types


class StatefulView(Stateful):
    def __init__(self, blocked_object: Stateful) -> None:
        check_argument_types()
        self._blocked_object = blocked_object

    @property
    def output(self):
        return self._blocked_object.output.detach()

    @property
    def dimension(self) -> int:
        return self._blocked_object.dimension

    @property
    def dependencies(self) -> List[str]:
        return super().dependencies + ["_blocked_object"]


class TemporalStatefulView(TemporalStateful):
    def __init__(self, blocked_object: TemporalStateful) -> None:
        check_argument_types()
        self._blocked_object = blocked_object

    @property
    def temporal_states(self):
        return self._blocked_object.temporal_states.detach()

    @property
    def temporal_mask(self):
        return self._blocked_object.temporal_mask

    @property
    def dimension(self) -> int:
        return self._blocked_object.dimension

    @property
    def dependencies(self) -> List[str]:
        return super().dependencies + ["_blocked_object"]


class SpatialStatefulView(SpatialStateful):
    def __init__(self, blocked_object: SpatialStateful) -> None:
        check_argument_types()
        self._blocked_object = blocked_object

    @property
    def spatial_states(self):
        return self._blocked_object.spatial_states.detach()

    @property
    def spatial_mask(self):
        return self._blocked_object.spatial_mask

    @property
    def dimension(self) -> int:
        return self._blocked_object.dimension

    @property
    def dependencies(self) -> List[str]:
        return super().dependencies + ["_blocked_object"]
