"""Greedy runner (reference: neuralmonkey/runners/runner.py:18-90).

The reference fetches `runtime_logprobs [T,B,V]` to the host and argmaxes there
(runner.py:49); with one session that is exactly the decoder's own greedy symbols, which
the fused logits kernel already produced on the device, so only `[T,B]` int64 crosses
PCIe.  With `num_sessions > 1` (`execute_sessions`) the reference's ensemble is reproduced as it is:
every session decodes on its own and the fetched log-probabilities are combined on the host.
"""
from typing import Any, Callable, Dict, List, Optional

import numpy as np

from neuralmonkey_b200.decoders.autoregressive import AutoregressiveDecoder
from neuralmonkey_b200.runners.base_runner import BaseRunner

Postprocessor = Optional[Callable[[List[List[str]]], List[List[str]]]]


class GreedyRunner(BaseRunner):
    class Executable(BaseRunner.Executable):
        def execute(self) -> None:
            runner = self.executor
            decoder = runner.decoder
            # argmax over the full vocabulary of the (single-session) log-probs == the greedy
            # feedback symbols without the `* unfinished` masking (runner.py:45-49)
            arg = decoder.runtime_argmax                      # [time, batch], computed on the device
            bsz = arg.shape[1]
            symbols = arg.cpu().numpy()
            train_loss = runtime_loss = 0.0
            if self.compute_losses:
                train_loss = float(decoder.train_loss)
                runtime_loss = float(decoder.runtime_loss)
            decoded_tokens = runner.vocabulary.vectors_to_sentences(symbols)
            if runner.postprocess is not None:
                decoded_tokens = runner.postprocess(decoded_tokens)
            self.set_runner_result(outputs=decoded_tokens, losses=[train_loss, runtime_loss],
                                   size=bsz)

        def execute_sessions(self, activate, num_sessions: int) -> None:
            """collect_results (runner.py:33-62) over several sessions: every session decodes greedily ON
            ITS OWN, the per-step log-probabilities [T, B, V] are fetched and combined with logaddexp
            step by step (over the steps of the first session), and the argmax of the sums is decoded;
            the losses are summed over the sessions."""
            runner = self.executor
            decoder = runner.decoder
            summed, train_loss, runtime_loss = None, 0.0, 0.0
            for index in range(num_sessions):
                activate(index)
                logprobs = decoder.runtime_logprobs.cpu().numpy()
                if self.compute_losses:
                    train_loss += float(decoder.train_loss)
                    runtime_loss += float(decoder.runtime_loss)
                if summed is None:
                    summed = [np.full(logprobs.shape[1:], -np.inf, dtype=logprobs.dtype) for _ in logprobs]
                for step, step_logprobs in enumerate(logprobs[:len(summed)]):   # the reference indexes past
                    summed[step] = np.logaddexp(summed[step], step_logprobs)     # the end if a later session is longer
            decoded_tokens = runner.vocabulary.vectors_to_sentences([np.argmax(s, axis=1) for s in summed])
            if runner.postprocess is not None:
                decoded_tokens = runner.postprocess(decoded_tokens)
            self.set_runner_result(outputs=decoded_tokens, losses=[train_loss, runtime_loss])

    def __init__(self, output_series: str, decoder: AutoregressiveDecoder,
                 postprocess: Postprocessor = None) -> None:
        BaseRunner.__init__(self, output_series, decoder)
        self.postprocess = postprocess
        self.vocabulary = self.decoder.vocabulary

    @property
    def loss_names(self) -> List[str]:
        return ["train_xent", "runtime_xent"]
