"""`AutoWriter`: picks the writer from the data (behaviour of neuralmonkey/writers/auto.py):
numpy array -> .npy; list of {name: array} -> .npz; list of token lists -> tokenized text; anything
else -> one `str(item)` per line."""
from typing import Any

import numpy as np

from neuralmonkey_b200.writers.numpy_writer import numpy_array_writer, numpy_dict_writer, stack_examples
from neuralmonkey_b200.writers.plain_text_writer import Writer, text_writer, tokenized_text_writer


def _is_dict_of_arrays(item: Any) -> bool:
    return isinstance(item, dict) and bool(item) and all(isinstance(k, str) for k in item)


def auto_writer(encoding: str = "utf-8") -> Writer:
    tokens_writer, plain_writer = tokenized_text_writer(encoding), text_writer(encoding)

    def writer(path: str, data: Any) -> None:
        if isinstance(data, np.ndarray):
            numpy_array_writer(path, data)
            return
        items = data if isinstance(data, list) else list(data)
        if items and all(_is_dict_of_arrays(item) for item in items):
            numpy_dict_writer(path, items)
        elif items and all(isinstance(item, (list, tuple)) and all(isinstance(t, str) for t in item)
                           for item in items):
            tokens_writer(path, items)
        elif items and not isinstance(items[0], str) and isinstance(items[0], (np.ndarray, list, tuple)):
            numpy_array_writer(path, stack_examples(items))     # per-example tensors
        else:
            plain_writer(path, items)
    return writer


# pylint: disable=invalid-name
AutoWriter = auto_writer()
