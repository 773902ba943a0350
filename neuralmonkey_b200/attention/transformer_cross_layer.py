"""Encoder-decoder attention combination for the Transformer decoder
(reference: neuralmonkey/attention/transformer_cross_layer.py:10-263): `single`, `serial`,
`parallel`, and the two further multi-source strategies `flat` and `hierarchical` (compositions of
`single`; GPU-verified by tests/test_gpu_variants.py - SURVEY.md 8(f) N4)."""
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
        return dropout(ctx, keep_prob, part.train_mode, residual=queries if residual else None)
    return ctx + queries if residual else ctx


def declare_cross(part, scope: str, strategy: str, dim: int, heads: List[int], heads_hier: int = None) -> None:
    from neuralmonkey_b200.encoders.transformer import declare_layer_norm
    if strategy == "flat":          # one attention over the concatenated encoders, in `scope` itself
        declare_single(part, scope, dim, heads[0], True)
        return
    if strategy in ("parallel", "hierarchical"):
        declare_layer_norm(part, scope, dim)
    for i, n_heads in enumerate(heads):
        declare_single(part, "{}/enc_{}".format(scope, i), dim, n_heads, strategy == "serial")
    if strategy == "hierarchical":
        declare_single(part, scope + "/enc_hier", dim, heads_hier, False)


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


def flat(part, scope: str, queries, encoder_states, encoder_masks, heads, attention_keep_probs,
         keep_prob) -> torch.Tensor:
    """States and masks concatenated along time, one attention over the lot (:228-263)."""
    return single(part, scope, queries, torch.cat(list(encoder_states), 1), torch.cat(list(encoder_masks), 1),
                  heads[0], attention_keep_probs[0], keep_prob)


def hierarchical(part, scope: str, queries, encoder_states, encoder_masks, heads, heads_hier,
                 attention_keep_probs, keep_prob) -> torch.Tensor:
    """Per-encoder contexts of the normalised queries, then a second attention of every query position
    over its own contexts ([batch*time, n_encoders, dim], all-ones mask, scope enc_hier), dropout,
    residual (:147-225)."""
    from neuralmonkey_b200.encoders.transformer import scoped_layer_norm
    normalized = scoped_layer_norm(part, scope, queries)
    contexts = [single(part, "{}/enc_{}".format(scope, i), normalized, states, mask, n_heads, akp, keep_prob,
                       normalize=False, residual=False)
                for i, (states, mask, n_heads, akp) in enumerate(zip(encoder_states, encoder_masks, heads,
                                                                     attention_keep_probs))]
    bsz, steps, dim = queries.shape
    stacked = torch.stack(contexts, 2).reshape(bsz * steps, len(contexts), dim)
    ones = torch.ones(bsz * steps, len(contexts), device=queries.device, dtype=torch.float32)
    ctx = single(part, scope + "/enc_hier", normalized.reshape(bsz * steps, 1, dim), stacked, ones, heads_hier,
                 keep_prob, 1.0, normalize=False, use_dropout=False, residual=False)
    return dropout(ctx.reshape(bsz, steps, dim), keep_prob, part.train_mode, residual=queries)
