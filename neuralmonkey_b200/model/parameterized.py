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
        if reuse is not None and initializers is not None:
            # the variables already exist in the other part's scope (parameterized.py:50-56)
            raise ValueError("Cannot use initializers in model part '{}' that reuses variables from '{}'."
                             .format(name, reuse.name))
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

    # -- per-part checkpoints (parameterized.py:109-125) ---------------------------------------
    def _scope_names(self) -> List[str]:
        """Names of the variables the part's Saver covers: `tf.get_collection(GLOBAL_VARIABLES, scope=name)`
        (parameterized.py:100-107) keeps every variable - trainable or not - whose name MATCHES the scope as a
        regular expression at its start, so `enc` also covers `enc_input/...` (a SentenceEncoder's checkpoint
        holds its input sequence's embeddings too)."""
        import re
        return [n for n in runtime.arena().order if re.match(self._scope, n)]

    def _scope_variables(self) -> Dict[str, torch.Tensor]:
        arena = runtime.arena()
        return {n: arena.get(n).detach().cpu().clone() for n in self._scope_names()}

    def save(self, session: Any = None) -> None:
        """Save the part's variables to its `save_checkpoint` file (no-op without one)."""
        if self._save_checkpoint:
            from neuralmonkey_b200.logging import log
            torch.save({"variables": self._scope_variables()}, self._save_checkpoint)
            log("Variables of '{}' saved to '{}'".format(self.name, self._save_checkpoint))

    def load(self, session: Any = None) -> None:
        """Load the part's variables from its `load_checkpoint` file (no-op without one).  The file is
        one written by `save` or a whole-model checkpoint of `TensorFlowManager.save`; variables of other
        scopes in it are ignored, a variable of this part missing from it is an error (Saver.restore)."""
        if self._load_checkpoint:
            from neuralmonkey_b200.logging import log
            ckpt = torch.load(self._load_checkpoint, map_location="cpu")
            stored = ckpt["variables"] if "variables" in ckpt else ckpt
            wanted = self._scope_names()
            missing = [n for n in wanted if n not in stored]
            if missing:
                raise KeyError("checkpoint '{}' holds no value for variable(s) {} of '{}'".format(
                    self._load_checkpoint, missing, self.name))
            runtime.arena().load_dict({n: stored[n] for n in wanted})
            log("Variables of '{}' loaded from '{}'".format(self.name, self._load_checkpoint))
