"""Vocabulary: word <-> index mapping (reference: neuralmonkey/vocabulary.py).

The reference converts strings to indices *inside* the TF graph with lookup tables
(vocabulary.py:187-195); here the lookup is a host-side dict and the result is an
int64 tensor, the dtype those tables emit.
"""
import json
import os
from typing import Dict, List, Optional, Sequence, Set, Union

import numpy as np
import torch

from neuralmonkey_b200.logging import log, notice, warn
from neuralmonkey_b200.typecheck import check_argument_types

PAD_TOKEN = "<pad>"
START_TOKEN = "<s>"
END_TOKEN = "</s>"
UNK_TOKEN = "<unk>"

SPECIAL_TOKENS = [PAD_TOKEN, START_TOKEN, END_TOKEN, UNK_TOKEN]

PAD_TOKEN_INDEX = 0
START_TOKEN_INDEX = 1
END_TOKEN_INDEX = 2
UNK_TOKEN_INDEX = 3


def from_wordlist(path: str, encoding: str = "utf-8", contains_header: bool = True,
                  contains_frequencies: bool = True) -> "Vocabulary":
    """Load a wordlist, optionally `word<TAB>count` with a header line (vocabulary.py:32-99).

    The four special tokens are expected on the first four lines and are not duplicated."""
    words = []  # type: List[str]
    with open(path, encoding=encoding) as wordlist:
        line_number = 1
        if contains_header:
            line_number += 1
            next(wordlist)
        for line in wordlist:
            line = line.strip()
            if not line:
                warn("Vocabulary file {}:{}: line empty".format(path, line_number))
                line_number += 1
                continue
            if contains_frequencies:
                info = line.split("\t")
                if len(info) != 2:
                    raise ValueError("Vocabulary file {}:{}: line does not have two columns"
                                     .format(path, line_number))
                word = info[0]
            else:
                if "\t" in line:
                    warn("Vocabulary file {}:{}: line contains a tabulator".format(path, line_number))
                word = line
            if line_number <= len(SPECIAL_TOKENS) + int(contains_header):
                should_be = SPECIAL_TOKENS[line_number - 1 - int(contains_header)]
                if word != should_be:
                    notice("Expected special token {} but encountered a different word: {}"
                           .format(should_be, word))
                    words.append(word)
                line_number += 1
                continue
            words.append(word)
            line_number += 1
    log("Vocabulary from wordlist loaded, containing {} words".format(len(words)))
    return Vocabulary(words)


def from_t2t_vocabulary(path: str, encoding: str = "utf-8") -> "Vocabulary":
    """A vocabulary file written by tensor2tensor (vocabulary.py:102-134): one entry per line, usually
    wrapped in a pair of single or double quotes (stripped); T2T's own `<pad>` and `<EOS>` entries are dropped,
    the special tokens of this toolkit take the first four indices as always."""
    check_argument_types()
    words = []  # type: List[str]
    with open(path, encoding=encoding) as wordlist:
        for line in wordlist:
            entry = line.strip()
            if entry and entry[0] == entry[-1] and entry[0] in "'\"":
                entry = entry[1:-1]
            if entry in ("<pad>", "<EOS>"):
                continue
            words.append(entry)
    log("Vocabulary form wordlist loaded, containing {} words".format(len(words)))
    return Vocabulary(words)


def from_nematus_json(path: str, max_size: int = None, pad_to_max_size: bool = False) -> "Vocabulary":
    """Nematus JSON vocabulary (vocabulary.py:137-172)."""
    with open(path, "r", encoding="utf-8") as f_json:
        contents = json.load(f_json)
    words = []  # type: List[str]
    for word in sorted(contents.keys(), key=lambda x: contents[x]):
        if contents[word] < 2:
            continue
        words.append(word)
        if max_size is not None and len(words) == max_size:
            break
    if max_size is None:
        max_size = len(words) - 2
    if pad_to_max_size and max_size is not None:
        current = len(words)
        for i in range(max_size - current + 2):
            words.append("<pad_{}>".format(i))
    return Vocabulary(words)


def from_bpe(path: str, encoding: str = "utf-8") -> "Vocabulary":
    """Compat shim for the stale `vocabulary.from_bpe` of examples/translation.ini:85: every
    symbol a BPE merge file can produce (both halves and the merged form of each rule)."""
    words = []  # type: List[str]
    seen = set()  # type: Set[str]
    with open(path, encoding=encoding) as f_bpe:
        for line in f_bpe:
            if line.startswith("#version"):
                continue
            parts = line.strip().split(" ")
            if len(parts) != 2:
                continue
            for sym in (parts[0], parts[1], parts[0] + parts[1]):
                for w in ((sym[:-4] if sym.endswith("</w>") else sym + "@@"),):
                    if w and w not in seen:
                        seen.add(w)
                        words.append(w)
    return Vocabulary(words)


class Vocabulary:
    def __init__(self, words: List[str], num_oov_buckets: int = 0) -> None:
        self._vocabulary = SPECIAL_TOKENS + list(words)
        self._alphabet = {c for word in words for c in word}
        self._word_to_index = {}  # type: Dict[str, int]
        for i, w in enumerate(self._vocabulary):
            self._word_to_index.setdefault(w, i)
        self.num_oov_buckets = num_oov_buckets

    def __len__(self) -> int:
        return len(self._vocabulary)

    def __contains__(self, word: str) -> bool:
        return word in self._word_to_index

    @property
    def alphabet(self) -> Set[str]:
        return self._alphabet

    @property
    def index_to_word(self) -> List[str]:
        return self._vocabulary

    def strings_to_indices(self, sentences: Sequence[Sequence[str]]) -> torch.Tensor:
        """[batch, time] token strings -> int64 CPU tensor; unknown words -> <unk>."""
        get = self._word_to_index.get
        arr = np.array([[get(w, UNK_TOKEN_INDEX) for w in sent] for sent in sentences],
                       dtype=np.int64)
        if arr.ndim == 1:
            arr = arr.reshape(len(sentences), 0)
        return torch.from_numpy(arr)

    def indices_to_strings(self, vectors) -> List[List[str]]:
        return [[self._vocabulary[int(i)] if 0 <= int(i) < len(self._vocabulary) else UNK_TOKEN
                 for i in row] for row in vectors]

    def vectors_to_sentences(self, vectors: Union[List[np.ndarray], np.ndarray]) -> List[List[str]]:
        """TIME-MAJOR index vectors -> token lists cut at </s> (vocabulary.py:257-288)."""
        if isinstance(vectors, list):
            if not vectors:
                raise ValueError("Cannot infer batch size because decoder returned an empty output.")
            batch_size = vectors[0].shape[0]
        elif isinstance(vectors, np.ndarray):
            batch_size = vectors.shape[1]
        else:
            raise TypeError("Unexpected type of decoder output: {}".format(type(vectors)))
        sentences = [[] for _ in range(batch_size)]  # type: List[List[str]]
        for vec in vectors:
            for sentence, word_i in zip(sentences, vec):
                if not sentence or sentence[-1] != END_TOKEN:
                    sentence.append(self.index_to_word[int(word_i)])
        return [s[:-1] if s and s[-1] == END_TOKEN else s for s in sentences]

    def save_wordlist(self, path: str, overwrite: bool = False, encoding: str = "utf-8") -> None:
        if os.path.exists(path) and not overwrite:
            raise FileExistsError("Cannot save vocabulary: File exists and overwrite is disabled. {}"
                                  .format(path))
        with open(path, "w", encoding=encoding) as output_file:
            for word in self._vocabulary:
                output_file.write("{}\n".format(word))


def pad_batch(sentences: List[List[str]], max_length: Optional[int] = None,
              add_start_symbol: bool = False, add_end_symbol: bool = False) -> List[List[str]]:
    """vocabulary.py:331-354: pad to the longest sentence (+1 for </s>), truncate to max_length."""
    max_len = max(len(s) for s in sentences)
    if add_end_symbol:
        max_len += 1
    if max_length is not None:
        max_len = min(max_length, max_len)
    padded_sentences = []
    for sent in sentences:
        if add_end_symbol:
            padded = (list(sent) + [END_TOKEN] + [PAD_TOKEN] * max_len)[:max_len]
        else:
            padded = (list(sent) + [PAD_TOKEN] * max_len)[:max_len]
        if add_start_symbol:
            padded.insert(0, START_TOKEN)
        padded_sentences.append(padded)
    return padded_sentences


def sentence_mask(indices: torch.Tensor) -> torch.Tensor:
    """float 0/1 mask of non-<pad> positions (vocabulary.py:357-358)."""
    return (indices != PAD_TOKEN_INDEX).to(torch.float32)
