"""Graph executors: trainers and runners (reference: neuralmonkey/runners/base_runner.py:21-195).

The reference's executables hand TF fetch dictionaries to `sess.run` and post-process the
numpy results (`next_to_execute` / `collect_results`).  Here there is no session: an
executable's `execute()` reads the lazily evaluated device tensors of its model part for
the batch currently fed, and `result` holds the same `ExecutionResult` structure.
"""
from typing import Any, Dict, List, NamedTuple, Optional, Set

from neuralmonkey_b200.model.feedable import Feedable
from neuralmonkey_b200.model.model_part import GenericModelPart
from neuralmonkey_b200.model.parameterized import Parameterized

ExecutionResult = NamedTuple("ExecutionResult", [
    ("outputs", Dict[str, Any]), ("losses", Dict[str, float]), ("size", int),
    ("summaries", List[Any])])


class GraphExecutor(GenericModelPart):
    class Executable:
        def __init__(self, executor: "GraphExecutor", compute_losses: bool, summaries: bool,
                     num_sessions: int) -> None:
            self._executor = executor
            self.compute_losses = compute_losses
            self.summaries = summaries
            self.num_sessions = num_sessions
            self._result = None  # type: Optional[ExecutionResult]

        def set_result(self, outputs: Dict[str, Any], losses: Dict[str, float], size: int,
                       summaries: List[Any]) -> None:
            self._result = ExecutionResult(outputs, losses, size, summaries)

        @property
        def result(self) -> Optional[ExecutionResult]:
            return self._result

        @property
        def executor(self):
            return self._executor

        def execute(self) -> None:
            """Compute and set the result for the batch currently fed."""
            raise NotImplementedError()

        def execute_sessions(self, activate, num_sessions: int) -> None:
            """Ensembles (`num_sessions > 1`): `activate(i)` switches the model parts to session i;
            run on every session and combine (the reference's next_to_execute / collect_results)."""
            raise NotImplementedError("{} cannot combine the outputs of several sessions".format(
                type(self.executor).__name__))

    def __init__(self, dependencies: Set[GenericModelPart]) -> None:
        self._dependencies = dependencies
        self._feedables, self._parameterizeds = self.get_dependencies()

    def get_executable(self, compute_losses: bool = False, summaries: bool = False,
                       num_sessions: int = 1):
        return self.Executable(self, compute_losses, summaries, num_sessions)

    @property
    def dependencies(self) -> List[str]:
        return ["_dependencies"]

    @property
    def feedables(self) -> Set[Feedable]:
        return self._feedables

    @property
    def parameterizeds(self) -> Set[Parameterized]:
        return self._parameterizeds


class BaseRunner(GraphExecutor):
    class Executable(GraphExecutor.Executable):
        def set_runner_result(self, outputs: Any, losses: List[float], size: int = None,
                              summaries: List[Any] = None) -> None:
            if summaries is None:
                summaries = []
            if size is None:
                size = len(outputs)
            loss_names = ["{}/{}".format(self.executor.output_series, loss)
                          for loss in self.executor.loss_names]
            self.set_result({self.executor.output_series: outputs}, dict(zip(loss_names, losses)),
                            size, summaries)

    def __init__(self, output_series: str, decoder: Any) -> None:
        GraphExecutor.__init__(self, {decoder})
        self.output_series = output_series
        self.decoder = decoder

    @property
    def decoder_data_id(self) -> Optional[str]:
        return getattr(self.decoder, "data_id", None)

    @property
    def loss_names(self) -> List[str]:
        raise NotImplementedError()
