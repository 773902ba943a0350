"""numpy writers of output series (names of neuralmonkey/writers/numpy_writer.py)."""
from typing import Any, Dict, Iterable, List

import numpy as np

from neuralmonkey_b200.logging import log


def stack_examples(items: List[Any]) -> np.ndarray:
    """A regular array when the examples agree in shape, an object array otherwise (bucketed
    batches give per-example tensors different time dimensions)."""
    try:
        return np.array(items)
    except ValueError:
        out = np.empty(len(items), dtype=object)
        for index, item in enumerate(items):
            out[index] = item
        return out


def numpy_array_writer(path: str, data: np.ndarray) -> None:
    np.save(path, data)
    log("Result saved as numpy array to '{}'".format(path))


def numpy_dict_writer(path: str, data: Iterable[Dict[str, np.ndarray]]) -> None:
    """A list of per-example {name: array} dicts -> one .npz with one stacked array per name."""
    examples = list(data)
    names = list(examples[0]) if examples else []
    np.savez(path, **{name: stack_examples([example[name] for example in examples]) for name in names})
    log("Result saved as numpy data to '{}.npz'".format(path))
