"""DatasetRunner: passes the input series through as outputs
(reference: neuralmonkey/runners/dataset_runner.py)."""
from neuralmonkey_b200.model.feedable import Feedable
from neuralmonkey_b200.runners.base_runner import GraphExecutor


class DatasetRunner(GraphExecutor, Feedable):
    class Executable(GraphExecutor.Executable):
        def execute(self) -> None:
            batch = self.executor.batch
            outputs = {s: list(batch.get_series(s)) for s in batch.series} if batch is not None else {}
            self.set_result(outputs, {}, len(batch) if batch is not None else 0, [])

        def execute_sessions(self, activate, num_sessions: int) -> None:
            self.execute()            # the input series do not depend on the session

    def __init__(self) -> None:
        GraphExecutor.__init__(self, set())
        Feedable.__init__(self)
        self.batch = None

    def feed_dict(self, dataset, train: bool = False):
        fd = Feedable.feed_dict(self, dataset, train)
        self.batch = dataset
        return fd

    def get_dependencies(self):
        return {self}, set()
