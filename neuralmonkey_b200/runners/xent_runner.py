"""XentRunner (reference: neuralmonkey/runners/xent_runner.py): the decoder's per-position training
cross-entropies `train_xents` [batch, time] as the output series - one list of floats per sentence - and
their mean over all entries as the loss "xent".  With several sessions the matrices are averaged."""
from typing import List

import numpy as np

from neuralmonkey_b200.decoders.autoregressive import AutoregressiveDecoder
from neuralmonkey_b200.runners.base_runner import BaseRunner
from neuralmonkey_b200.typecheck import check_argument_types


class XentRunner(BaseRunner):
    class Executable(BaseRunner.Executable):
        def _finish(self, xents: np.ndarray) -> None:
            self.set_runner_result(outputs=xents.tolist(), losses=[float(np.mean(xents))])

        def execute(self) -> None:
            self._finish(self.executor.decoder.train_xents.detach().cpu().numpy())

        def execute_sessions(self, activate, num_sessions: int) -> None:
            per_session = []
            for index in range(num_sessions):
                activate(index)
                per_session.append(self.executor.decoder.train_xents.detach().cpu().numpy())
            self._finish(np.mean(per_session, axis=0))

    def __init__(self, output_series: str, decoder: AutoregressiveDecoder) -> None:
        check_argument_types()
        BaseRunner.__init__(self, output_series, decoder)

    @property
    def loss_names(self) -> List[str]:
        return ["xent"]
