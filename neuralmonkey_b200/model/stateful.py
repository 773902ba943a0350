"""Stateful interfaces (reference: neuralmonkey/model/stateful.py:22-103)."""
from neuralmonkey_b200.model.model_part import GenericModelPart


class Stateful(GenericModelPart):
    @property
    def output(self):
        """[batch, state_size] tensor."""
        raise NotImplementedError("Abstract property")


class TemporalStateful(GenericModelPart):
    @property
    def temporal_states(self):
        """[batch, time, state_size] tensor."""
        raise NotImplementedError("Abstract property")

    @property
    def temporal_mask(self):
        """[batch, time] float 0/1 tensor."""
        raise NotImplementedError("Abstract property")

    @property
    def lengths(self):
        """int32 [batch]: sum of the mask (stateful.py:56-62)."""
        return self.temporal_mask.sum(dim=1).to(dtype=__import__("torch").int32)

    @property
    def dimension(self) -> int:
        raise NotImplementedError("Abstract property")


class SpatialStateful(GenericModelPart):
    @property
    def spatial_states(self):
        raise NotImplementedError("Abstract property")

    @property
    def spatial_mask(self):
        raise NotImplementedError("Abstract property")

    @property
    def dimension(self) -> int:
        raise NotImplementedError("Abstract property")


class TemporalStatefulWithOutput(Stateful, TemporalStateful):
    pass


class SpatialStatefulWithOutput(Stateful, SpatialStateful):
    pass
