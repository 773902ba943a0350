"""Attention loop states (reference: neuralmonkey/attention/namedtuples.py)."""
from typing import List, NamedTuple

import torch

AttentionLoopState = NamedTuple("AttentionLoopState",
                                [("contexts", torch.Tensor), ("weights", torch.Tensor)])
HierarchicalLoopState = NamedTuple("HierarchicalLoopState",
                                   [("child_loop_states", List), ("loop_state", AttentionLoopState)])
MultiHeadLoopState = NamedTuple("MultiHeadLoopState",
                                [("contexts", torch.Tensor), ("head_weights", List[torch.Tensor])])
