"""Beam search runner (reference: neuralmonkey/runners/beamsearch_runner.py:14-190):
picks the `rank`-th hypothesis of the final beam, drops the start slot, cuts at </s>."""
from typing import Callable, List

import numpy as np

from neuralmonkey_b200.decoders.beam_search_decoder import BeamSearchDecoder
from neuralmonkey_b200.runners.base_runner import BaseRunner
from neuralmonkey_b200.vocabulary import END_TOKEN_INDEX


class BeamSearchRunner(BaseRunner):
    class Executable(BaseRunner.Executable):
        def execute(self) -> None:
            runner = self.executor
            if self.num_sessions != 1:
                raise NotImplementedError("beam search ensembles (num_sessions > 1) are not built")
            output = runner.decoder.outputs.last_search_step_output
            scores = output.scores.cpu().numpy()
            bs_scores = [s[runner.rank - 1] for s in scores]
            tok_ids = np.transpose(output.token_ids.cpu().numpy(), [1, 2, 0])
            index_to_word = runner.decoder.vocabulary.index_to_word
            decoded_tokens = []
            for toks in tok_ids:
                decoded = []
                for tok_id in toks[runner.rank - 1][1:]:
                    if tok_id == END_TOKEN_INDEX:
                        break
                    decoded.append(index_to_word[tok_id])
                decoded_tokens.append(decoded)
            if runner.postprocess is not None:
                decoded_tokens = runner.postprocess(decoded_tokens)
            self.set_runner_result(outputs=decoded_tokens,
                                   losses=[float(np.mean(bs_scores) * len(bs_scores))])

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
