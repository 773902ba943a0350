"""Multi-head scaled dot-product attention
(reference: neuralmonkey/attention/scaled_dot_product.py:24-226).

`attention()` = q/k/v projections (bias-free by default) -> K8 core (`ops.mha_core`:
scaling, causal and key masks with the reference's -1e9 semantics, softmax, PV) -> output
projection.  With one head the reference applies no projections at all (:171-179,217-223).
Variables are declared by the owning model part under `<scope>/{query,keys,vals,output}_proj`.
"""
from typing import Optional, Tuple

import torch

from neuralmonkey_b200 import ops
from neuralmonkey_b200.nn.utils import dropout, dropout_mask
from neuralmonkey_b200.nn.variants import require_variant
from neuralmonkey_b200.params import zeros_initializer


def declare_attention(part, scope: str, q_dim: int, kv_dim: int, num_heads: int,
                      use_bias: bool = False) -> None:
    """Variables of one `attention()` call site (tf.layers.dense names)."""
    if num_heads <= 1:
        return
    for name, in_dim in (("query_proj", q_dim), ("keys_proj", kv_dim), ("vals_proj", kv_dim),
                         ("output_proj", q_dim)):
        part.declare("{}/{}/kernel".format(scope, name), [in_dim, q_dim])
        if use_bias:
            part.declare("{}/{}/bias".format(scope, name), [q_dim], zeros_initializer())


def attention(part, scope: str, queries: torch.Tensor, keys: torch.Tensor, values: torch.Tensor,
              keys_mask: Optional[torch.Tensor], num_heads: int, masked: bool = False,
              attention_dropout_keep_prob: float = 1.0, train_mode: bool = False,
              use_bias: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (context [batch, time_q, dim], weights [batch, heads, time_q, time_k])."""
    if num_heads <= 0:
        raise ValueError("Number of heads must be greater than zero.")
    q_dim = queries.shape[-1]
    if q_dim != keys.shape[-1]:
        raise ValueError("Queries and keys do not match in the last dimension. Queries: {}, Keys: {}"
                         .format(q_dim, keys.shape[-1]))
    if keys.shape[1] != values.shape[1]:
        raise ValueError("Keys and values 'time' dimension does not match. Keys: {}, Values: {}"
                         .format(keys.shape[1], values.shape[1]))
    if q_dim % num_heads != 0:
        raise ValueError("Last dimension of the query ({}) should be divisible by the number of "
                         "heads ({})".format(q_dim, num_heads))

    def proj(x, name):
        bias = part.var("{}/{}/bias".format(scope, name)) if use_bias else None
        return ops.linear(x, part.var("{}/{}/kernel".format(scope, name)), bias)

    if num_heads > 1:
        queries, keys, values = proj(queries, "query_proj"), proj(keys, "keys_proj"), proj(values, "vals_proj")
    drop = None
    if attention_dropout_keep_prob < 1.0 and train_mode:
        # dropout on the attention weights (:208-214), inside the fused core: the mask rides along
        require_variant("attention_dropout_keep_prob < 1")
        drop = dropout_mask((queries.shape[0], num_heads, queries.shape[1], keys.shape[1]),
                            attention_dropout_keep_prob, train_mode, queries.device)
    context, weights = ops.mha_core(queries, keys, values, keys_mask, masked, num_heads, drop)
    if num_heads > 1:
        context = proj(context, "output_proj")
    return context, weights
