"""Dropout helper (reference: neuralmonkey/nn/utils.py:6-22).

`tf.nn.dropout` selected by `train_mode`.  The random mask is drawn with torch's CUDA
generator (host plumbing; the streams cannot match TF's, so parity tests run with
keep_prob = 1 or eval mode) and applied as one elementwise product.
"""
from typing import Optional

import torch


def dropout_mask(shape, keep_prob: float, train_mode: bool, device) -> Optional[torch.Tensor]:
    """Mask already scaled by 1/keep_prob, or None when dropout is inactive."""
    if keep_prob >= 1.0 or not train_mode:
        return None
    mask = (torch.rand(shape, device=device) < keep_prob).to(torch.float32)
    return mask / keep_prob


def dropout(variable: torch.Tensor, keep_prob: float, train_mode: bool) -> torch.Tensor:
    mask = dropout_mask(variable.shape, keep_prob, train_mode, variable.device)
    if mask is None:
        return variable
    return variable * mask
