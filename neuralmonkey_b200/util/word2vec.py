"""word2vec text files as a vocabulary plus an embedding initialiser
(reference: neuralmonkey/util/word2vec.py; tests/language-model.ini).

File format: a header line `<number of words> <dimension>`, then one `word v1 ... vD` line per word.  The
matrix starts with one row per special token (<pad>, <s>, </s>, <unk> - zeros unless the file has a vector for
that token), followed by the file's other words in file order: the row order `Vocabulary(words)` assigns."""
from typing import Callable, List, Sequence

import numpy as np
import torch

from neuralmonkey_b200.typecheck import check_argument_types
from neuralmonkey_b200.vocabulary import SPECIAL_TOKENS, Vocabulary


class Word2Vec:
    def __init__(self, path: str, encoding: str = "utf-8") -> None:
        check_argument_types()
        words = []  # type: List[str]
        with open(path, encoding=encoding) as lines:
            dimension = int(next(lines).split()[1])
            rows = [np.zeros(dimension) for _ in SPECIAL_TOKENS]
            for line in lines:
                fields = line.split()
                vector = np.array([float(x) for x in fields[1:]], dtype=np.float64)
                assert vector.shape[0] == dimension
                if fields[0] in SPECIAL_TOKENS:
                    rows[SPECIAL_TOKENS.index(fields[0])] = vector
                else:
                    words.append(fields[0])
                    rows.append(vector)
        self.vocab = Vocabulary(words)
        self.embedding_matrix = np.stack(rows)

    @property
    def vocabulary(self) -> Vocabulary:
        return self.vocab

    @property
    def embeddings(self) -> np.ndarray:
        return self.embedding_matrix


def get_word2vec_initializer(w2v: Word2Vec) -> Callable:
    """An initialiser for `initializers=[("word_embeddings", <...>)]` that returns the file's matrix; a
    variable of any other shape is an error."""
    check_argument_types()

    def init(shape: Sequence[int], *_args, **_kwargs) -> torch.Tensor:
        if list(shape) != list(w2v.embeddings.shape):
            raise ValueError("Shapes of model and word2vec embeddings do not match. Word2Vec shape: {}, "
                             "Should have been: {}".format(w2v.embeddings.shape, list(shape)))
        return torch.from_numpy(np.ascontiguousarray(w2v.embeddings)).to(torch.float64)

    return init


def word2vec_vocabulary(w2v: Word2Vec) -> Vocabulary:
    check_argument_types()
    return w2v.vocabulary
