"""Gradient-blocking views of encoder-like objects (reference: neuralmonkey/model/gradient_blocking.py;
tests/bpe.ini freezes its encoder this way in the second run of tests_run.sh): the same states with the
graph cut behind them, so nothing upstream receives a gradient through the view.

A view is not a model part - it owns no variables and is fed nothing - but the object it wraps stays reachable
through `dependencies`, so its variables and feeds are still collected.  The values are per-batch tensors of the
wrapped object: a view re-reads them at every access (`detach()` is a view of the same memory, nothing is
copied or cached), masks and the state width are passed through as they are."""
from typing import Any, List

from neuralmonkey_b200.model.stateful import SpatialStateful, Stateful, TemporalStateful
from neuralmonkey_b200.typecheck import check_type


def _cut(attribute: str) -> property:
    """The wrapped object's tensor attribute, detached from the graph."""
    return property(lambda view: getattr(view._blocked_object, attribute).detach())  # pylint: disable=protected-access


def _same(attribute: str) -> property:
    return property(lambda view: getattr(view._blocked_object, attribute))  # pylint: disable=protected-access


class _BlockingView:
    """What the three views share: the wrapped object (type-checked against the interface the view stands for)
    and its place among the dependencies."""
    _interface = object  # type: Any

    def __init__(self, blocked_object) -> None:
        check_type("blocked_object", blocked_object, self._interface)
        self._blocked_object = blocked_object

    dimension = _same("dimension")

    @property
    def dependencies(self) -> List[str]:
        return super().dependencies + ["_blocked_object"]


class StatefulView(_BlockingView, Stateful):
    _interface = Stateful
    output = _cut("output")


class TemporalStatefulView(_BlockingView, TemporalStateful):
    _interface = TemporalStateful
    temporal_states = _cut("temporal_states")
    temporal_mask = _same("temporal_mask")


class SpatialStatefulView(_BlockingView, SpatialStateful):
    _interface = SpatialStateful
    spatial_states = _cut("spatial_states")
    spatial_mask = _same("spatial_mask")
