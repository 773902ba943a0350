"""Writers of output series (reference: neuralmonkey/writers/plain_text_writer.py, writers/auto.py).
A writer is called as writer(path, data)."""
from typing import Any, Callable, Iterable, List

import numpy as np

Writer = Callable[[str, Any], None]


def text_writer(encoding: str = "utf-8") -> Writer:
    def writer(path: str, data: Iterable[List[str]]) -> None:
        with open(path, "w", encoding=encoding) as f_out:
            for sentence in data:
                f_out.write(" ".join(str(tok) for tok in sentence) + "\n")
    return writer


def numpy_writer(path: str, data: Any) -> None:
    np.save(path, np.asarray(list(data), dtype=object) if not isinstance(data, np.ndarray) else data)


def AutoWriter(path: str, data: Any) -> None:  # pylint: disable=invalid-name
    """Text for lists of token lists / strings, numpy otherwise (writers/auto.py:20-60)."""
    data = list(data) if not isinstance(data, (list, np.ndarray)) else data
    if isinstance(data, np.ndarray):
        np.save(path, data)
        return
    if all(isinstance(item, str) for item in data):
        with open(path, "w", encoding="utf-8") as f_out:
            f_out.write("\n".join(data) + "\n")
        return
    if all(isinstance(item, (list, tuple)) and all(isinstance(t, str) for t in item) for item in data):
        text_writer()(path, data)
        return
    if all(isinstance(item, dict) for item in data):
        np.savez(path, **{k: _stack([d[k] for d in data]) for k in data[0]}) if data else np.savez(path)
        return
    np.save(path, _stack(data))


def _stack(items: List[Any]) -> np.ndarray:
    """Regular array when the examples agree in shape, object array otherwise (bucketed
    batches give per-example tensors different time dimensions)."""
    try:
        return np.array(items)
    except ValueError:
        out = np.empty(len(items), dtype=object)
        for i, item in enumerate(items):
            out[i] = item
        return out


UtfPlainTextWriter = text_writer()
