"""Edit-operation series for post-editing models (reference: neuralmonkey/processors/editops.py;
tests/post-edit.ini): a dataset-level preprocessor that turns (machine translation, post-edited text) into
the sequence of operations leading from one to the other - `<keep>`, `<delete>`, or the word to insert - and
the postprocessor that applies a decoded operation sequence to the translation again.

The operations are those of a cheapest keep / delete / insert script (no substitutions: a keep costs nothing
and needs equal words, the other two cost 1).  Among equally cheap scripts the reference's table prefers, cell
by cell, keep over delete over insert; the same choice is recorded here as a back pointer per cell and the
script is read off backwards."""
from typing import Any, Callable, Dict, Iterable, Iterator, List

KEEP = "<keep>"
DELETE = "<delete>"

_KEEP, _DELETE, _INSERT = 0, 1, 2


def convert_to_edits(source: List[str], target: List[str]) -> List[str]:
    n_src, n_tgt = len(source), len(target)
    # cost[i][j]: cheapest script turning source[:i] into target[:j]; move[i][j]: its last operation
    cost = [[0] * (n_tgt + 1) for _ in range(n_src + 1)]
    move = [[_INSERT] * (n_tgt + 1) for _ in range(n_src + 1)]
    for i in range(1, n_src + 1):
        cost[i][0], move[i][0] = i, _DELETE
    for j in range(1, n_tgt + 1):
        cost[0][j] = j
    for i in range(1, n_src + 1):
        row, above = cost[i], cost[i - 1]
        for j in range(1, n_tgt + 1):
            best, how = above[j] + 1, _DELETE
            if source[i - 1] == target[j - 1] and above[j - 1] <= best:
                best, how = above[j - 1], _KEEP
            if row[j - 1] + 1 < best:
                best, how = row[j - 1] + 1, _INSERT
            row[j], move[i][j] = best, how
    script = []  # type: List[str]
    i, j = n_src, n_tgt
    while i > 0 or j > 0:
        how = move[i][j]
        if how == _KEEP:
            script.append(KEEP)
            i, j = i - 1, j - 1
        elif how == _DELETE:
            script.append(DELETE)
            i -= 1
        else:
            script.append(target[j - 1])
            j -= 1
    script.reverse()
    return script


def reconstruct(source: List[str], edits: List[str]) -> List[str]:
    """Apply an operation sequence.  A keep beyond the end of the source yields nothing; source words the
    script never reached (the decoder stopped early) are copied."""
    position = 0
    out = []  # type: List[str]
    for op in edits:
        if op == KEEP:
            if position < len(source):
                out.append(source[position])
            position += 1
        elif op == DELETE:
            position += 1
        else:
            out.append(op)
    out.extend(source[position:])
    return out


# pylint: disable=too-few-public-methods
class Preprocess:
    """Dataset-level preprocessor: the edit operations from series `source_id` to series `target_id`."""

    def __init__(self, source_id: str, target_id: str) -> None:
        self._source_id = source_id
        self._target_id = target_id

    def __call__(self, iterators: Dict[str, Callable[[], Iterator[List[str]]]]) -> Iterator[List[str]]:
        for src_seq, tgt_seq in zip(iterators[self._source_id](), iterators[self._target_id]()):
            yield convert_to_edits(src_seq, tgt_seq)


class Postprocess:
    """Postprocessor: the generated series `edits_id` applied to the dataset series `source_id`."""

    def __init__(self, source_id: str, edits_id: str) -> None:
        self._source_id = source_id
        self._edits_id = edits_id

    def __call__(self, dataset: Dict[str, Iterable[Any]], generated: Dict[str, Iterable[Any]]) -> List[List[str]]:
        if self._source_id not in dataset:
            raise ValueError("Source series not present in the input dataset")
        if self._edits_id not in generated:
            raise ValueError("Edits series not present in the output dataset")
        return [reconstruct(src, ops) for src, ops in zip(dataset[self._source_id], generated[self._edits_id])]
