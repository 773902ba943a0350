"""Initial-state projections (reference: neuralmonkey/decoders/encoder_projection.py:37-145).

An EncoderProjection is a callable (decoder, train_mode, rnn_size, encoders) -> [batch, rnn_size]
plus a `declare(decoder, rnn_size, encoders)` hook naming its variables under the decoder's
`initial_state/` scope.
"""
from typing import Callable, List, Optional

import torch

from neuralmonkey_b200 import ops, runtime
from neuralmonkey_b200.model.stateful import Stateful
from neuralmonkey_b200.nn.utils import dropout
from neuralmonkey_b200.params import zeros_initializer


class EncoderProjection:
    def declare(self, decoder, rnn_size: Optional[int], encoders: List[Stateful]) -> None:
        pass

    def output_size(self, rnn_size: Optional[int], encoders: List[Stateful]) -> int:
        raise NotImplementedError

    def __call__(self, decoder, train_mode: bool, rnn_size: Optional[int],
                 encoders: List[Stateful]) -> torch.Tensor:
        raise NotImplementedError


def _encoder_output_size(enc: Stateful) -> int:
    return enc.dimension


class _Empty(EncoderProjection):
    def output_size(self, rnn_size, encoders):
        if rnn_size is None:
            raise ValueError("You must supply rnn_size for this type of encoder projection")
        return rnn_size

    def __call__(self, decoder, train_mode, rnn_size, encoders):
        if rnn_size is None:
            raise ValueError("You must supply rnn_size for this type of encoder projection")
        return torch.zeros(rnn_size, device=runtime.device())


class _Concat(EncoderProjection):
    def output_size(self, rnn_size, encoders):
        if not encoders:
            raise ValueError("There must be at least one encoder for this type of encoder projection")
        size = sum(_encoder_output_size(e) for e in encoders)
        if rnn_size is not None and rnn_size != size:
            raise ValueError("RNN size supplied for concat projection ({}) does not match the size "
                             "of the concatenated vectors ({}).".format(rnn_size, size))
        return size

    def __call__(self, decoder, train_mode, rnn_size, encoders):
        self.output_size(rnn_size, encoders)
        outs = [e.output for e in encoders]
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)


class _Linear(EncoderProjection):
    def __init__(self, dropout_keep_prob: float) -> None:
        self.dropout_keep_prob = dropout_keep_prob

    def output_size(self, rnn_size, encoders):
        if rnn_size is None:
            raise ValueError("You must supply rnn_size for this type of encoder projection")
        return rnn_size

    def declare(self, decoder, rnn_size, encoders):
        in_size = sum(_encoder_output_size(e) for e in encoders)
        decoder.declare("initial_state/encoders_projection/kernel", [in_size, rnn_size])
        decoder.declare("initial_state/encoders_projection/bias", [rnn_size], zeros_initializer())

    def __call__(self, decoder, train_mode, rnn_size, encoders):
        en_concat = concat_encoder_projection(decoder, train_mode, None, encoders)
        y = ops.linear(en_concat, decoder.var("initial_state/encoders_projection/kernel"),
                       decoder.var("initial_state/encoders_projection/bias"))
        return dropout(y, self.dropout_keep_prob, train_mode)


empty_initial_state = _Empty()
concat_encoder_projection = _Concat()


def linear_encoder_projection(dropout_keep_prob: float) -> EncoderProjection:
    """dropout(dense(concat(encoder outputs), rnn_size)) (encoder_projection.py:47-73)."""
    return _Linear(dropout_keep_prob)
