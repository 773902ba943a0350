"""Beam search runner (reference: neuralmonkey/runners/beamsearch_runner.py:14-190):
picks the `rank`-th hypothesis of the final beam, drops the start slot, cuts at </s>."""
from typing import Callable, List, Tuple

import numpy as np

from neuralmonkey_b200.decoders.beam_search_decoder import BeamSearchDecoder
from neuralmonkey_b200.runners.base_runner import BaseRunner
from neuralmonkey_b200.vocabulary import END_TOKEN_INDEX


def select_hypotheses(scores: np.ndarray, token_ids: np.ndarray, rank: int,
                      index_to_word: List[str]) -> Tuple[List[List[str]], float]:
    """Host post-processing of a finished search (beamsearch_runner.py:81-103, `prepare_results`).

    `scores` is [batch, beam] (best first), `token_ids` is [time, batch, beam] with the start symbol
    in slot 0.  Returns the words of every sentence's `rank`-th hypothesis up to (excluding) the
    first </s>, and the runner's loss: the sum of the selected hypotheses' scores.

    One deliberate difference: a hypothesis that is empty (its first symbol is </s>) comes out as
    `[]`.  The reference assigns the word list inside the token loop (:94), so for an empty
    hypothesis it leaves the raw array of token ids in the output batch; the test documents this."""
    bs_scores = [s[rank - 1] for s in scores]
    decoded_tokens = []
    for toks in np.transpose(token_ids, [1, 2, 0]):
        decoded = []
        for tok_id in toks[rank - 1][1:]:
            if tok_id == END_TOKEN_INDEX:
                break
            decoded.append(index_to_word[tok_id])
        decoded_tokens.append(decoded)
    return decoded_tokens, float(np.mean(bs_scores) * len(bs_scores))


class BeamSearchRunner(BaseRunner):
    class Executable(BaseRunner.Executable):
        def execute(self) -> None:
            runner = self.executor
            output = runner.decoder.outputs.last_search_step_output
            decoded_tokens, loss = select_hypotheses(
                output.scores.cpu().numpy(), output.token_ids.cpu().numpy(), runner.rank,
                runner.decoder.vocabulary.index_to_word)
            if runner.postprocess is not None:
                decoded_tokens = runner.postprocess(decoded_tokens)
            self.set_runner_result(outputs=decoded_tokens, losses=[loss])

        def execute_sessions(self, activate, num_sessions: int) -> None:
            """Ensembled search (beamsearch_runner.py:44-118): one beam, every session steps its own
            decoder on it, the next-token log-probabilities are averaged in probability space."""
            runner = self.executor
            output = runner.decoder.ensemble_outputs(activate, num_sessions).last_search_step_output
            decoded_tokens, loss = select_hypotheses(
                output.scores.cpu().numpy(), output.token_ids.cpu().numpy(), runner.rank,
                runner.decoder.vocabulary.index_to_word)
            if runner.postprocess is not None:
                decoded_tokens = runner.postprocess(decoded_tokens)
            self.set_runner_result(outputs=decoded_tokens, losses=[loss])

    def __init__(self, output_series: str, decoder: BeamSearchDecoder, rank: int = 1,
                 postprocess: Callable[[List[str]], List[str]] = None) -> None:
        BaseRunner.__init__(self, output_series, decoder)
        if rank < 1 or rank > decoder.beam_size:
            raise ValueError("Rank of output hypothesis must be between 1 and the beam size ({}), "
                             "was {}.".format(decoder.beam_size, rank))
        self.rank = rank
        self.postprocess = postprocess

    @property
    def loss_names(self) -> List[str]:
        return ["beam_search_score"]


def beam_search_runner_range(output_series: str, decoder: BeamSearchDecoder, max_rank: int = None,
                             postprocess: Callable[[List[str]], List[str]] = None
                             ) -> List[BeamSearchRunner]:
    """Runners for ranks 1..max_rank; series names `<output_series>.rankNNN`."""
    if max_rank is None:
        max_rank = decoder.beam_size
    if max_rank > decoder.beam_size:
        raise ValueError("The maximum rank ({}) cannot be bigger than beam size {}.".format(
            max_rank, decoder.beam_size))
    return [BeamSearchRunner("{}.rank{:03d}".format(output_series, r), decoder, r, postprocess)
            for r in range(1, max_rank + 1)]
