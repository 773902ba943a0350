"""Transformer encoder (reference: neuralmonkey/encoders/transformer.py:23-330):
pre-LayerNorm blocks x = x + drop(MHA(LN(x))), x = x + drop(FFN(LN(x))), a final LayerNorm,
sinusoidal position signal with concatenated (not interleaved) sin / cos halves, and
`output` = unmasked sum over time (:170-172)."""
import math
from typing import List, NamedTuple

import torch

from neuralmonkey_b200.typecheck import check_argument_types
from neuralmonkey_b200 import ops, runtime
from neuralmonkey_b200.attention.base_attention import (Attendable, get_attention_mask,
                                                        get_attention_states)
from neuralmonkey_b200.attention.scaled_dot_product import attention, declare_attention
from neuralmonkey_b200.decorators import tensor
from neuralmonkey_b200.model.model_part import ModelPart
from neuralmonkey_b200.model.parameterized import InitializerSpecs
from neuralmonkey_b200.model.stateful import TemporalStateful, TemporalStatefulWithOutput
from neuralmonkey_b200.nn.utils import dropout
from neuralmonkey_b200.params import (ones_initializer, variance_scaling_initializer,
                                      zeros_initializer)

TransformerLayer = NamedTuple("TransformerLayer", [("temporal_states", torch.Tensor),
                                                   ("temporal_mask", torch.Tensor)])

_position_cache = {}


def position_signal(dimension: int, length: int) -> torch.Tensor:
    """[1, length, dimension] timing signal (transformer.py:23-45), fp32 arithmetic."""
    key = (dimension, length, str(runtime.device()))
    if key not in _position_cache:
        positions = torch.arange(length, dtype=torch.float32)
        num_timescales = dimension // 2
        log_increment = math.log(1.0e4) / (num_timescales - 1)
        inv_timescales = torch.exp(torch.arange(num_timescales, dtype=torch.float32) * -log_increment)
        scaled_time = positions.unsqueeze(1) * inv_timescales.unsqueeze(0)
        signal = torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1)
        if dimension % 2:
            signal = torch.nn.functional.pad(signal, (0, 1))
        _position_cache[key] = signal.reshape(1, length, dimension).to(runtime.device())
    return _position_cache[key]


def declare_layer_norm(part, scope: str, dim: int) -> None:
    part.declare(scope + "/LayerNorm/gamma", [dim], ones_initializer())
    part.declare(scope + "/LayerNorm/beta", [dim], zeros_initializer())


def scoped_layer_norm(part, scope: str, x: torch.Tensor) -> torch.Tensor:
    prefix = scope + "/LayerNorm/" if scope else "LayerNorm/"
    return ops.layer_norm(x, part.var(prefix + "gamma"), part.var(prefix + "beta"))


def declare_feedforward(part, scope: str, dim: int, hidden: int) -> None:
    declare_layer_norm(part, scope, dim)
    part.declare(scope + "/hidden_state/kernel", [dim, hidden])
    part.declare(scope + "/hidden_state/bias", [hidden], zeros_initializer())
    part.declare(scope + "/output/kernel", [hidden, dim])
    part.declare(scope + "/output/bias", [dim], zeros_initializer())


def feedforward_sublayer(part, scope: str, layer_input: torch.Tensor, keep_prob: float,
                         train_mode: bool) -> torch.Tensor:
    """dense-relu-drop-dense-drop + residual on LN(x) (transformer.py:266-288)."""
    normalized = scoped_layer_norm(part, scope, layer_input)
    hidden = ops.linear(normalized, part.var(scope + "/hidden_state/kernel"),
                        part.var(scope + "/hidden_state/bias"), act="relu")
    hidden = dropout(hidden, keep_prob, train_mode)
    out = ops.linear(hidden, part.var(scope + "/output/kernel"), part.var(scope + "/output/bias"))
    return dropout(out, keep_prob, train_mode, residual=layer_input)


class TransformerEncoder(ModelPart, TemporalStatefulWithOutput):
    # pylint: disable=too-many-arguments,too-many-instance-attributes
    def __init__(self, name: str, input_sequence: TemporalStateful, ff_hidden_size: int, depth: int,
                 n_heads: int, dropout_keep_prob: float = 1.0,
                 attention_dropout_keep_prob: float = 1.0, target_space_id: int = None,
                 use_att_transform_bias: bool = False, use_positional_encoding: bool = True,
                 input_for_cross_attention: Attendable = None, n_cross_att_heads: int = None,
                 reuse: ModelPart = None, save_checkpoint: str = None, load_checkpoint: str = None,
                 initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.input_sequence = input_sequence
        self.ff_hidden_size = ff_hidden_size
        self.depth = depth
        self.n_heads = n_heads
        self.dropout_keep_prob = dropout_keep_prob
        self.attention_dropout_keep_prob = attention_dropout_keep_prob
        self.target_space_id = target_space_id
        self.use_att_transform_bias = use_att_transform_bias
        self.use_positional_encoding = use_positional_encoding
        self.input_for_cross_attention = input_for_cross_attention
        self.n_cross_att_heads = n_cross_att_heads
        if self.depth <= 0:
            raise ValueError("Depth must be a positive integer.")
        if self.ff_hidden_size <= 0:
            raise ValueError("Feed forward hidden size must be a positive integer.")
        if self.dropout_keep_prob <= 0.0 or self.dropout_keep_prob > 1.0:
            raise ValueError("Dropout keep prob must be inside (0,1].")
        if self.attention_dropout_keep_prob <= 0.0 or self.attention_dropout_keep_prob > 1.0:
            raise ValueError("Dropout keep prob for attn must be in (0,1].")
        if self.target_space_id is not None and (self.target_space_id >= 32 or self.target_space_id < 0):
            raise ValueError("If provided, the target space ID should be between 0 and 31. Was: {}"
                             .format(self.target_space_id))
        if (input_for_cross_attention is None) != (n_cross_att_heads is None):
            raise ValueError("Either both input_for_cross_attention and n_cross_att_heads must be "
                             "provided or none of them.")
        self._default_initializer = variance_scaling_initializer(mode="fan_avg", distribution="uniform")

    @property
    def model_dimension(self) -> int:
        dim = self.input_sequence.dimension
        if self.input_for_cross_attention is not None:
            if self.input_for_cross_attention.dimension != dim:
                raise ValueError("The input for cross-attention must be of the same dimension as "
                                 "the model, was {}.".format(self.input_for_cross_attention.dimension))
        return dim

    @property
    def dimension(self) -> int:
        return self.model_dimension

    @property
    def dependencies(self) -> List[str]:
        deps = ModelPart.dependencies.fget(self)
        return deps + ["input_for_cross_attention"] if self.input_for_cross_attention is not None else deps

    def declare_variables(self) -> None:
        if hasattr(self.input_sequence, "ensure_declared"):
            self.input_sequence.ensure_declared()
        dim = self.model_dimension
        if self.target_space_id is not None:
            self.declare("target_modality_embedding_matrix", [32, dim])
        for i in range(self.depth):
            scope = "layer_{}".format(i)
            declare_layer_norm(self, scope + "/self_attention", dim)
            declare_attention(self, scope + "/self_attention", dim, dim, self.n_heads,
                              self.use_att_transform_bias)
            if self.input_for_cross_attention is not None:
                declare_layer_norm(self, scope + "/cross_attention", dim)
                declare_attention(self, scope + "/cross_attention", dim, dim, self.n_cross_att_heads,
                                  self.use_att_transform_bias)
            declare_feedforward(self, scope + "/feedforward", dim, self.ff_hidden_size)
        self.declare("LayerNorm/gamma", [dim], ones_initializer())
        self.declare("LayerNorm/beta", [dim], zeros_initializer())

    @tensor
    def encoder_inputs(self) -> torch.Tensor:
        inputs = self.input_sequence.temporal_states
        if self.target_space_id is not None:
            inputs = inputs + self.var("target_modality_embedding_matrix")[self.target_space_id].view(1, 1, -1)
        if self.use_positional_encoding:
            inputs = inputs + position_signal(self.model_dimension, inputs.shape[1])
        return dropout(inputs, self.dropout_keep_prob, self.train_mode)

    def _layer(self, i: int, states: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        scope = "layer_{}".format(i)
        normalized = scoped_layer_norm(self, scope + "/self_attention", states)
        ctx, _ = attention(self, scope + "/self_attention", normalized, normalized, normalized, mask,
                           self.n_heads, False, self.attention_dropout_keep_prob, self.train_mode,
                           self.use_att_transform_bias)
        states = dropout(ctx, self.dropout_keep_prob, self.train_mode, residual=states)
        if self.input_for_cross_attention is not None:
            enc_states = get_attention_states(self.input_for_cross_attention)
            enc_mask = get_attention_mask(self.input_for_cross_attention)
            normalized = scoped_layer_norm(self, scope + "/cross_attention", states)
            ctx, _ = attention(self, scope + "/cross_attention", normalized, enc_states, enc_states,
                               enc_mask, self.n_cross_att_heads, False,
                               self.attention_dropout_keep_prob, self.train_mode,
                               self.use_att_transform_bias)
            states = dropout(ctx, self.dropout_keep_prob, self.train_mode, residual=states)
        return feedforward_sublayer(self, scope + "/feedforward", states, self.dropout_keep_prob,
                                    self.train_mode)

    @tensor
    def temporal_states(self) -> torch.Tensor:
        states, mask = self.encoder_inputs, self.temporal_mask
        for i in range(self.depth):
            states = self._layer(i, states, mask)
        return scoped_layer_norm(self, "", states)

    @tensor
    def temporal_mask(self) -> torch.Tensor:
        return self.input_sequence.temporal_mask

    @tensor
    def output(self) -> torch.Tensor:
        return self.temporal_states.sum(dim=1)
