"""Byte-pair-encoding pre- and post-processors
(reference: neuralmonkey/processors/bpe.py:10-60, which wraps the vendored subword-nmt
`apply_bpe`; algorithm: Sennrich, Haddow & Birch 2016, arxiv.org/abs/1508.07909).

The merge table is a priority list of symbol pairs (earlier line = higher priority, the first
occurrence of a duplicate wins).  A word is split into characters plus an end-of-word marker;
repeatedly, the adjacent pair with the best priority is merged everywhere in the word until no
pair of the table is left.  Pieces are emitted with the separator appended to all but the last.
"""
import re
from typing import Dict, List, Tuple

from neuralmonkey_b200.logging import log

END_OF_WORD = "</w>"


class BPEPreprocessor:
    def __init__(self, merge_file: str, separator: str = "@@", encoding: str = "utf-8") -> None:
        log("Initializing BPE preprocessor")
        self.separator = separator
        self.ranks = {}  # type: Dict[Tuple[str, ...], int]
        with open(merge_file, "r", encoding=encoding) as f_data:
            for rank, line in enumerate(f_data):
                pair = tuple(line.split())
                if pair not in self.ranks:
                    self.ranks[pair] = rank
        self._cache = {}  # type: Dict[str, Tuple[str, ...]]

    def _segment_word(self, word: str) -> Tuple[str, ...]:
        if word in self._cache:
            return self._cache[word]
        symbols = list(word) + [END_OF_WORD]
        while len(symbols) > 1:
            best_rank, best_pair = None, None
            for pair in zip(symbols, symbols[1:]):
                rank = self.ranks.get(pair)
                if rank is not None and (best_rank is None or rank < best_rank):
                    best_rank, best_pair = rank, pair
            if best_pair is None:
                break
            merged, i = [], 0
            while i < len(symbols):
                if i + 1 < len(symbols) and (symbols[i], symbols[i + 1]) == best_pair:
                    merged.append(symbols[i] + symbols[i + 1])
                    i += 2
                else:
                    merged.append(symbols[i])
                    i += 1
            symbols = merged
        if symbols[-1] == END_OF_WORD:
            symbols = symbols[:-1]
        elif symbols[-1].endswith(END_OF_WORD):
            symbols[-1] = symbols[-1].replace(END_OF_WORD, "")
        result = tuple(symbols)
        self._cache[word] = result
        return result

    def __call__(self, sentence: List[str]) -> List[str]:
        output = []
        for word in sentence:
            if not word:
                output.append(word)
                continue
            pieces = self._segment_word(word)
            output.extend(piece + self.separator for piece in pieces[:-1])
            output.append(pieces[-1])
        return output


class BPEPostprocessor:
    def __init__(self, separator: str = "@@") -> None:
        self.pattern = re.compile(re.escape(separator) + r" ")

    def __call__(self, decoded_sentences: List[List[str]]) -> List[List[str]]:
        return [self.decode(s) for s in decoded_sentences]

    def decode(self, sentence: List[str]) -> List[str]:
        return self.pattern.sub("", " ".join(sentence)).split(" ")
