"""Dropout helper (reference: neuralmonkey/nn/utils.py:6-22).

`tf.nn.dropout` selected by `train_mode`.  On the GPU the keep decisions are drawn INSIDE the kernel that applies
them (`ops.dropout`, K15: Philox counters keyed by the experiment's seed, the training step and the call's number
within the step), so a dropout is one launch forward and one backward and no mask tensor exists; kernels that take a
mask operand (attention weights, the decoder's recurrent state) get theirs from `ops.dropout_mask`.  The random
streams cannot match TensorFlow's, so parity tests run with keep_prob = 1, in eval mode, or with a deterministic
mask patched in for `dropout_mask` - a patched or CPU-side mask is applied as a plain product.
"""
from typing import Optional

import torch

from neuralmonkey_b200 import ops


def dropout_mask(shape, keep_prob: float, train_mode: bool, device) -> Optional[torch.Tensor]:
    """Mask already scaled by 1/keep_prob, or None when dropout is inactive."""
    if keep_prob >= 1.0 or not train_mode:
        return None
    if torch.device(device).type == "cuda":
        return ops.dropout_mask(shape, keep_prob, device)
    mask = (torch.rand(shape, device=device) < keep_prob).to(torch.float32)
    return mask / keep_prob


_BUILTIN_MASK = dropout_mask


def dropout(variable: torch.Tensor, keep_prob: float, train_mode: bool,
            residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dropout(variable) [+ residual]: the residual connection of the Transformer sublayers rides in the same pass."""
    if keep_prob != 1.0 and not 0.0 < keep_prob <= 1.0:
        # tf.nn.dropout's check; the reference builds the op whatever the mode, so this is raised in eval mode too
        raise ValueError("keep_prob must be a scalar tensor or a float in the range (0, 1], got {:g}".format(keep_prob))
    if keep_prob >= 1.0 or not train_mode:
        return variable if residual is None else variable + residual
    if variable.is_cuda and dropout_mask is _BUILTIN_MASK:
        return ops.dropout(variable, keep_prob, residual)
    mask = dropout_mask(variable.shape, keep_prob, train_mode, variable.device)
    out = variable if mask is None else variable * mask
    return out if residual is None else out + residual
