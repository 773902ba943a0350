"""Readers of pre-computed numpy arrays (reference: neuralmonkey/readers/numpy_reader.py:8-47): the
series a `SpatialFiller` / `StatefulFiller` consumes."""
import os
from typing import Callable, Iterable, List

import numpy as np

from neuralmonkey_b200.typecheck import check_argument_types


def single_tensor(files: List[str]) -> np.ndarray:
    """One array holding the whole series; several files are joined along the first axis."""
    check_argument_types()
    arrays = [np.load(path) for path in files]
    return arrays[0] if len(arrays) == 1 else np.concatenate(arrays, axis=0)


def from_file_list(prefix: str, shape: List[int], suffix: str = "",
                   default_tensor_name: str = "arr_0") -> Callable:
    """Each line of the list files names an .npz under `prefix`; yields its `default_tensor_name`."""
    check_argument_types()

    def load(files: List[str]) -> Iterable[np.ndarray]:
        for list_file in files:
            with open(list_file, encoding="utf-8") as f_list:
                for line in f_list:
                    path = os.path.join(prefix, line.rstrip()) + suffix
                    with np.load(path) as npz:
                        arr = npz[default_tensor_name]
                    if list(arr.shape) != shape:
                        raise ValueError("Shapes do not match: expected {}, found {}".format(
                            shape, list(arr.shape)))
                    yield arr
    return load
