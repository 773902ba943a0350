"""TensorRunner / RepresentationRunner: dump arbitrary model-part attributes
(reference: neuralmonkey/runners/tensor_runner.py:14-204).  The parity hook: any lazily
evaluated tensor of a model part can be written out by name."""
from typing import Dict, List

import numpy as np

from neuralmonkey_b200.model.model_part import GenericModelPart
from neuralmonkey_b200.runners.base_runner import BaseRunner


class TensorRunner(BaseRunner):
    class Executable(BaseRunner.Executable):
        def _fetch(self) -> List:
            """The tensors of the active session as a batch of per-instance values."""
            runner = self.executor
            fetched = {}
            for part, name, bdim in zip(runner.modelparts, runner.tensor_names, runner.batch_dims):
                value = getattr(part, name).detach().float().cpu().numpy()
                if bdim != 0:
                    value = np.moveaxis(value, bdim, 0)
                fetched[name] = value
            if len(fetched) == 1 and runner.single_tensor:
                return list(next(iter(fetched.values())))
            n = len(next(iter(fetched.values())))
            return [{k: v[i] for k, v in fetched.items()} for i in range(n)]

        def execute(self) -> None:
            outputs = self._fetch()
            self.set_runner_result(outputs=outputs, losses=[], size=len(outputs))

        def execute_sessions(self, activate, num_sessions: int) -> None:
            """tensor_runner.py:26-40: with `select_session` the tensors of that session, otherwise every
            instance gets the tuple of its values in all sessions."""
            select = self.executor.select_session
            if select is not None:
                activate(select)
                outputs = self._fetch()
            else:
                sessions = []
                for index in range(num_sessions):
                    activate(index)
                    sessions.append(self._fetch())
                outputs = list(zip(*sessions))
            self.set_runner_result(outputs=outputs, losses=[], size=len(outputs))

    # pylint: disable=too-many-arguments
    def __init__(self, output_series: str, modelparts: List[GenericModelPart], tensors: List[str],
                 batch_dims: List[int], tensors_by_name: List[str] = None,
                 batch_dims_by_name: List[int] = None, select_session: int = None,
                 single_tensor: bool = False) -> None:
        if not modelparts:
            raise ValueError("At least one model part is expected")
        BaseRunner.__init__(self, output_series, modelparts[0])
        if len(modelparts) != len(tensors) or len(tensors) != len(batch_dims):
            raise ValueError("modelparts, tensors and batch_dims must have the same length")
        if tensors_by_name:
            raise NotImplementedError("tensors_by_name refers to TF graph names; use attributes")
        for part, name in zip(modelparts, tensors):
            if not hasattr(type(part), name) and not hasattr(part, name):
                raise TypeError("The model part {} does not have a tensor called {}.".format(part, name))
        self.modelparts = modelparts
        self.tensor_names = tensors
        self.batch_dims = batch_dims
        self.single_tensor = single_tensor
        self.select_session = select_session
        self._dependencies = set(modelparts)
        self._feedables, self._parameterizeds = self.get_dependencies()

    @property
    def loss_names(self) -> List[str]:
        return []


class RepresentationRunner(TensorRunner):
    """Dump one attribute (default `output`) of an encoder (tensor_runner.py:160-204)."""

    def __init__(self, output_series: str, encoder: GenericModelPart, attribute: str = "output",
                 select_session: int = None) -> None:
        TensorRunner.__init__(self, output_series, modelparts=[encoder], tensors=[attribute],
                              batch_dims=[0], single_tensor=True)
