"""Series of pre-computed numpy arrays - what a `SpatialFiller` / `StatefulFiller` consumes.
Same entry points and error behaviour as neuralmonkey/readers/numpy_reader.py:8-47 of the reference."""
import os
from typing import Callable, Iterator, List

import numpy as np

from neuralmonkey_b200.typecheck import check_argument_types


def single_tensor(files: List[str]) -> np.ndarray:
    """The whole series stored as ONE array (first axis = instance); several files are joined on that axis."""
    check_argument_types()
    parts = list(map(np.load, files))
    if len(parts) > 1:
        return np.concatenate(parts, axis=0)
    return parts[0]


class _ArchiveList:
    """Reader over list files: every line names an .npz archive `<prefix>/<line><suffix>`."""

    def __init__(self, prefix: str, shape: List[int], suffix: str, key: str) -> None:
        self.prefix, self.shape, self.suffix, self.key = prefix, list(shape), suffix, key

    def _member(self, entry: str) -> np.ndarray:
        with np.load(os.path.join(self.prefix, entry) + self.suffix) as archive:
            array = archive[self.key]
        found = list(array.shape)
        if found != self.shape:
            raise ValueError("Shapes do not match: expected {}, found {}".format(self.shape, found))
        return array

    def __call__(self, files: List[str]) -> Iterator[np.ndarray]:
        for listing in files:
            with open(listing, encoding="utf-8") as entries:
                yield from (self._member(entry.rstrip()) for entry in entries)


def from_file_list(prefix: str, shape: List[int], suffix: str = "",
                   default_tensor_name: str = "arr_0") -> Callable:
    check_argument_types()
    return _ArchiveList(prefix, shape, suffix, default_tensor_name)
