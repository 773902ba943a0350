"""Encoder-decoder attention combination for the Transformer decoder
(reference: neuralmonkey/attention/transformer_cross_layer.py:10-147): `single`, `serial`
and `parallel`.  `flat` and `hierarchical` are multi-source strategies outside the hot path
(one encoder)."""
from typing import List

import torch

from neuralmonkey_b200.attention.scaled_dot_product import attention, declare_attention
from neuralmonkey_b200.nn.utils import dropout


def declare_single(part, scope: str, dim: int, n_heads: int, normalize: bool = True) -> None:
    from neuralmonkey_b200.encoders.transformer import declare_layer_norm
    if normalize:
        declare_layer_norm(part, scope, dim)
    declare_attention(part, scope, dim, dim, n_heads, False)


def single(part, scope: str, queries: torch.Tensor, states: torch.Tensor, mask: torch.Tensor,
           n_heads: int, attention_keep_prob: float, keep_prob: float, normalize: bool = True,
           use_dropout: bool = True, residual: bool = True) -> torch.Tensor:
    from neuralmonkey_b200.encoders.transformer import scoped_layer_norm
    normalized = scoped_layer_norm(part, scope, queries) if normalize else queries
    ctx, _ = attention(part, scope, normalized, states, states, mask, n_heads, False,
                       attention_keep_prob, part.train_mode, False)
    if use_dropout:
        ctx = dropout(ctx, keep_prob, part.train_mode)
    if residual:
        ctx = ctx + queries
    return ctx


def declare_cross(part, scope: str, strategy: str, dim: int, heads: List[int]) -> None:
    from neuralmonkey_b200.encoders.transformer import declare_layer_norm
    if strategy == "parallel":
        declare_layer_norm(part, scope, dim)
    for i, n_heads in enumerate(heads):
        declare_single(part, "{}/enc_{}".format(scope, i), dim, n_heads, strategy == "serial")


def serial(part, scope: str, queries, encoder_states, encoder_masks, heads, attention_keep_probs,
           keep_prob) -> torch.Tensor:
    context = queries
    for i, (states, mask, n_heads, akp) in enumerate(zip(encoder_states, encoder_masks, heads,
                                                         attention_keep_probs)):
        context = single(part, "{}/enc_{}".format(scope, i), context, states, mask, n_heads, akp,
                         keep_prob)
    return context


def parallel(part, scope: str, queries, encoder_states, encoder_masks, heads, attention_keep_probs,
             keep_prob) -> torch.Tensor:
    from neuralmonkey_b200.encoders.transformer import scoped_layer_norm
    normalized = scoped_layer_norm(part, scope, queries)
    total = queries
    for i, (states, mask, n_heads, akp) in enumerate(zip(encoder_states, encoder_masks, heads,
                                                         attention_keep_probs)):
        total = total + single(part, "{}/enc_{}".format(scope, i), normalized, states, mask, n_heads,
                               akp, keep_prob, normalize=False, residual=False)
    return total
