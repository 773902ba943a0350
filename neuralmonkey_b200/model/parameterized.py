"""Parameterized: owns named variables in the flat arena
(reference: neuralmonkey/model/parameterized.py:15-125).

Variable names are `<part name>/<local name>`, the reference's variable-scope naming;
`reuse=<other part>` shares the other part's scope (parameterized.py:44-60);
`initializers=[(local name, initializer)]` overrides per variable (tf_utils.py:35-51).
"""
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from neuralmonkey_b200 import runtime
from neuralmonkey_b200.params import Initializer, normal_initializer

InitializerSpecs = List[Tuple[str, Any]]


class Parameterized:
    def __init__(self, name: str, reuse: "Parameterized" = None, save_checkpoint: str = None,
                 load_checkpoint: str = None, initializers: InitializerSpecs = None) -> None:
        self._name = name
        self._reuse = reuse
        self._save_checkpoint = save_checkpoint
        self._load_checkpoint = load_checkpoint
        self._initializers = dict(initializers) if initializers else {}  # type: Dict[str, Any]
        self._default_initializer = normal_initializer(stddev=0.001)
        self._declared = False
        self._scope = reuse.scope_name if reuse is not None else name

    @property
    def name(self) -> str:
        return self._name

    @property
    def scope_name(self) -> str:
        return self._scope

    def __str__(self) -> str:
        return "{}: {}".format(type(self).__name__, self._name)

    # -- variable declaration / access ------------------------------------------------
    def declare(self, local_name: str, shape: Sequence[int], initializer: Optional[Initializer] = None,
                trainable: bool = True, absolute: bool = False) -> None:
        full = local_name if absolute else "{}/{}".format(self._scope, local_name)
        init = self._initializers.get(local_name, initializer or self._default_initializer)
        runtime.arena().declare(full, shape, init, trainable)

    def var(self, local_name: str, absolute: bool = False) -> torch.Tensor:
        full = local_name if absolute else "{}/{}".format(self._scope, local_name)
        return runtime.arena().get(full)

    def declare_variables(self) -> None:
        """Declare every variable of this part (called once before the arena is finalized)."""

    def ensure_declared(self) -> None:
        if not self._declared:
            self._declared = True
            self.declare_variables()
