"""Numeric vectors stored as text, one whitespace-separated vector per non-empty line (optionally gzipped).
Same entry points as neuralmonkey/readers/string_vector_reader.py:6-38 of the reference."""
import gzip
from typing import Iterator, List, Optional, Type

import numpy as np


class _VectorLines:
    def __init__(self, dtype: Type, columns: Optional[int]) -> None:
        self.dtype, self.columns = dtype, columns

    def __call__(self, files: List[str]) -> Iterator[np.ndarray]:
        for path in files:
            handle = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
            with handle:
                for lineno, text in enumerate(handle, start=1):
                    fields = text.split()
                    if not fields:
                        continue
                    if self.columns is not None and len(fields) != self.columns:
                        raise ValueError("Wrong number of columns ({}) on line {}, file {}".format(
                            len(fields), lineno, path))
                    yield np.array(fields, dtype=self.dtype)


def get_string_vector_reader(dtype: Type = np.float32, columns: int = None):
    return _VectorLines(dtype, columns)


# pylint: disable=invalid-name
FloatVectorReader = get_string_vector_reader(np.float32)
IntVectorReader = get_string_vector_reader(np.int32)
