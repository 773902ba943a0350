"""Attention base class and encoder-state helpers
(reference: neuralmonkey/attention/base_attention.py:54-205)."""
from typing import Any, Dict, Optional, Tuple, Union

import torch

from neuralmonkey_b200 import runtime
from neuralmonkey_b200.attention.namedtuples import AttentionLoopState
from neuralmonkey_b200.model.model_part import ModelPart
from neuralmonkey_b200.model.parameterized import InitializerSpecs
from neuralmonkey_b200.model.stateful import SpatialStateful, TemporalStateful

Attendable = Union[TemporalStateful, SpatialStateful]


def empty_attention_loop_state(batch_size: int, length: int, dimension: int) -> AttentionLoopState:
    dev = runtime.device()
    return AttentionLoopState(contexts=torch.zeros(0, batch_size, dimension, device=dev),
                              weights=torch.zeros(0, batch_size, length, device=dev))


def get_attention_states(encoder: Attendable) -> torch.Tensor:
    """[batch, time, dim]: temporal states, or spatial states flattened over (h, w)
    (base_attention.py:79-97)."""
    if isinstance(encoder, TemporalStateful):
        return encoder.temporal_states
    if isinstance(encoder, SpatialStateful):
        s = encoder.spatial_states
        return s.reshape(s.shape[0], s.shape[1] * s.shape[2], s.shape[3])
    raise TypeError("Unknown encoder type")


def get_attention_mask(encoder: Attendable) -> Optional[torch.Tensor]:
    if isinstance(encoder, TemporalStateful):
        if encoder.temporal_mask is None:
            raise ValueError("The encoder temporal mask should not be none")
        return encoder.temporal_mask
    if isinstance(encoder, SpatialStateful):
        if encoder.spatial_mask is None:
            return None
        m = encoder.spatial_mask
        return m.reshape(m.shape[0], m.shape[1] * m.shape[2])
    raise TypeError("Unknown encoder type")


class BaseAttention(ModelPart):
    def __init__(self, name: str, reuse: ModelPart = None, save_checkpoint: str = None,
                 load_checkpoint: str = None, initializers: InitializerSpecs = None) -> None:
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.query_state_size = None  # type: Optional[int]
        self._histories = {}  # type: Dict[str, torch.Tensor]

    @property
    def histories(self) -> Dict[str, torch.Tensor]:
        return self._histories

    def attention(self, query: torch.Tensor, decoder_prev_state: torch.Tensor,
                  decoder_input: torch.Tensor, loop_state: Any) -> Tuple[torch.Tensor, Any]:
        raise NotImplementedError("Abstract method")

    def record_weights(self, key: str, weights: torch.Tensor) -> None:
        """Keep the weights of a whole pass, as `attention_sequence` returned them (batch-major,
        [batch, NQ, time]), under `key` - time-major, like the histories of a stepped loop."""
        self.histories[key] = weights.detach().transpose(0, 1)

    def initial_loop_state(self) -> Any:
        raise NotImplementedError("Abstract method")

    def finalize_loop(self, key: str, last_loop_state: Any) -> None:
        raise NotImplementedError("Abstract method")

    @property
    def context_vector_size(self) -> int:
        raise NotImplementedError("Abstract property")

    def visualize_attention(self, key: str, max_outputs: int = 16) -> None:
        """TensorBoard image summaries of the reference (base_attention.py:187-205): a no-op
        here beyond the key check; the histories stay available in `self.histories`."""
        if key not in self.histories:
            raise KeyError("Key {} not among attention histories".format(key))
