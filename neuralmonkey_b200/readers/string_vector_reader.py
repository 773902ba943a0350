"""Vectors written as whitespace-separated numbers, one per line
(reference: neuralmonkey/readers/string_vector_reader.py:6-46)."""
import gzip
from typing import Iterable, List, Type

import numpy as np


def get_string_vector_reader(dtype: Type = np.float32, columns: int = None):
    def parse(line: str, lineno: int, path: str) -> np.ndarray:
        numbers = line.split()
        if columns is not None and len(numbers) != columns:
            raise ValueError("Wrong number of columns ({}) on line {}, file {}".format(
                len(numbers), lineno, path))
        return np.array(numbers, dtype=dtype)

    def reader(files: List[str]) -> Iterable[np.ndarray]:
        for path in files:
            opener = (lambda p: gzip.open(p, "rt")) if path.endswith(".gz") else open
            with opener(path) as f_data:
                for lineno, line in enumerate(f_data, 1):
                    if line.strip():
                        yield parse(line, lineno, path)
    return reader


# pylint: disable=invalid-name
FloatVectorReader = get_string_vector_reader(np.float32)
IntVectorReader = get_string_vector_reader(np.int32)
