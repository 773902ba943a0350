"""Sequence model parts (reference: neuralmonkey/model/sequence.py:54-300).

`EmbeddedSequence` turns one data series of token strings into a length-masked sequence of
embeddings.  The gather + mask is kernel K1 (`ops.embed`).
"""
from typing import Any, Dict, List, Optional

import torch

from neuralmonkey_b200.typecheck import check_argument_types
from neuralmonkey_b200 import ops, runtime
from neuralmonkey_b200.decorators import tensor
from neuralmonkey_b200.model.model_part import ModelPart
from neuralmonkey_b200.model.parameterized import InitializerSpecs
from neuralmonkey_b200.model.stateful import TemporalStateful
from neuralmonkey_b200.vocabulary import Vocabulary, pad_batch, sentence_mask


class Sequence(ModelPart, TemporalStateful):
    def __init__(self, name: str, max_length: int = None, reuse: ModelPart = None,
                 save_checkpoint: str = None, load_checkpoint: str = None,
                 initializers: InitializerSpecs = None) -> None:
        ModelPart.__init__(self, name, reuse, save_checkpoint, load_checkpoint, initializers)
        self.max_length = max_length
        if self.max_length is not None and self.max_length <= 0:
            raise ValueError("Max sequence length must be a positive integer.")


class EmbeddedFactorSequence(Sequence):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, vocabularies: List[Vocabulary], data_ids: List[str],
                 embedding_sizes: List[int], max_length: int = None, add_start_symbol: bool = False,
                 add_end_symbol: bool = False, scale_embeddings_by_depth: bool = False,
                 trainable: bool = True, embeddings_source: "EmbeddedFactorSequence" = None,
                 reuse: ModelPart = None, save_checkpoint: str = None, load_checkpoint: str = None,
                 initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        Sequence.__init__(self, name, max_length, reuse, save_checkpoint, load_checkpoint,
                          initializers)
        self.vocabularies = vocabularies
        self.vocabulary_sizes = [len(vocab) for vocab in self.vocabularies]
        self.data_ids = data_ids
        self.embedding_sizes = embedding_sizes
        self.add_start_symbol = add_start_symbol
        self.add_end_symbol = add_end_symbol
        self.scale_embeddings_by_depth = scale_embeddings_by_depth
        self.embeddings_source = embeddings_source
        self.trainable = trainable
        if not (len(self.data_ids) == len(self.vocabularies) == len(self.embedding_sizes)):
            raise ValueError("data_ids, vocabularies, and embedding_sizes lists need to have "
                             "the same length")
        if any(es <= 0 for es in self.embedding_sizes):
            raise ValueError("Embedding size must be a positive integer.")
        if embeddings_source is not None:
            if list(self.vocabularies) != list(embeddings_source.vocabularies):
                raise ValueError("When reusing embeedings, vocabularies must be the same.")
            if list(self.embedding_sizes) != list(embeddings_source.embedding_sizes):
                raise ValueError("When reusing embeedings, embeddings sizes must be equal.")
        self._ids = []  # type: List[torch.Tensor]

    def declare_variables(self) -> None:
        if self.embeddings_source is not None:
            self.embeddings_source.ensure_declared()
            return
        for i, (vsize, esize) in enumerate(zip(self.vocabulary_sizes, self.embedding_sizes)):
            self.declare("embedding_matrix_{}".format(i), [vsize, esize], trainable=self.trainable)

    @property
    def embedding_matrices(self) -> List[torch.Tensor]:
        if self.embeddings_source is not None:
            return self.embeddings_source.embedding_matrices
        return [self.var("embedding_matrix_{}".format(i)) for i in range(len(self.data_ids))]

    @property
    def input_types(self) -> Dict[str, Any]:
        return {d_id: str for d_id in self.data_ids}

    @property
    def input_shapes(self) -> Dict[str, Any]:
        return {d_id: [None, None] for d_id in self.data_ids}

    # -- feeding ---------------------------------------------------------------------
    def feed_dict(self, dataset, train: bool = False) -> Dict[str, Any]:
        fd = ModelPart.feed_dict(self, dataset, train)
        ids = []
        for vocab, name in zip(self.vocabularies, self.data_ids):
            sentences = dataset.get_series(name)
            padded = pad_batch(list(sentences), self.max_length, self.add_start_symbol,
                               self.add_end_symbol)
            ids.append(vocab.strings_to_indices(padded))
            fd[name] = ids[-1]
        self._set_ids(ids)
        return fd

    def feed_ids(self, ids: List[torch.Tensor], train: bool = False) -> None:
        """Feed already-indexed factors ([batch, time] int64), bypassing strings."""
        self.reset_batch()
        self.train_mode = bool(train)
        self.batch_size = int(ids[0].shape[0])
        self._set_ids(ids)

    def _set_ids(self, ids: List[torch.Tensor]) -> None:
        dev = runtime.device()
        self._ids = [runtime.to_device(i) for i in ids]

    def static_inputs(self) -> Dict[str, Any]:
        return {"ids{}".format(i): t for i, t in enumerate(self._ids)}

    def bind_static(self, tensors: Dict[str, Any]) -> None:
        self.reset_batch()
        self._ids = [tensors["ids{}".format(i)] for i in range(len(self._ids))]

    @tensor
    def input_factor_indices(self) -> List[torch.Tensor]:
        return self._ids

    @tensor
    def temporal_mask(self) -> torch.Tensor:
        return sentence_mask(self.input_factor_indices[0])

    @tensor
    def temporal_states(self) -> torch.Tensor:
        """Embedded factors * mask, concatenated on the feature axis (sequence.py:170-194)."""
        mask = self.temporal_mask
        factors = []
        for ids, matrix in zip(self.input_factor_indices, self.embedding_matrices):
            m = mask
            if self.scale_embeddings_by_depth:
                m = mask * (matrix.shape[-1] ** 0.5)
            factors.append(ops.embed(ids, matrix, m))
        return factors[0] if len(factors) == 1 else torch.cat(factors, 2)

    @property
    def dimension(self) -> int:
        return sum(self.embedding_sizes)


class EmbeddedSequence(EmbeddedFactorSequence):
    # pylint: disable=too-many-arguments
    def __init__(self, name: str, vocabulary: Vocabulary, data_id: str, embedding_size: int,
                 max_length: int = None, add_start_symbol: bool = False,
                 add_end_symbol: bool = False, scale_embeddings_by_depth: bool = False,
                 trainable: bool = True, embeddings_source: "EmbeddedSequence" = None,
                 reuse: ModelPart = None, save_checkpoint: str = None, load_checkpoint: str = None,
                 initializers: InitializerSpecs = None) -> None:
        check_argument_types()
        EmbeddedFactorSequence.__init__(
            self, name=name, vocabularies=[vocabulary], data_ids=[data_id],
            embedding_sizes=[embedding_size], max_length=max_length,
            add_start_symbol=add_start_symbol, add_end_symbol=add_end_symbol,
            scale_embeddings_by_depth=scale_embeddings_by_depth, trainable=trainable,
            embeddings_source=embeddings_source, reuse=reuse, save_checkpoint=save_checkpoint,
            load_checkpoint=load_checkpoint, initializers=initializers)

    @property
    def inputs(self) -> torch.Tensor:
        return self.input_factor_indices[0]

    @property
    def embedding_matrix(self) -> torch.Tensor:
        return self.embedding_matrices[0]

    @property
    def vocabulary(self) -> Vocabulary:
        return self.vocabularies[0]

    @property
    def data_id(self) -> str:
        return self.data_ids[0]
