"""Text writers of output series (names and behaviour of neuralmonkey/writers/plain_text_writer.py).
A writer is called as `writer(path, data)`; `tokenized_text_writer` inverts `tokenized_text_reader`,
`t2t_tokenized_text_writer` inverts `t2t_tokenized_text_reader`."""
from typing import Any, Callable, Iterable, Iterator, List

from neuralmonkey_b200.logging import log
from neuralmonkey_b200.readers.plain_text_reader import _is_alnum

Writer = Callable[[str, Any], None]


def t2t_detokenize(data: Iterable[List[str]]) -> Iterator[str]:
    """Glue tensor2tensor-style tokens back together: a space goes only between two neighbouring
    tokens that both start with an alphanumeric character."""
    for sentence in data:
        pieces, previous_alnum = [], False
        for position, token in enumerate(sentence):
            starts_alnum = _is_alnum(token[0])
            if position and previous_alnum and starts_alnum:
                pieces.append(" ")
            pieces.append(token)
            previous_alnum = starts_alnum
        yield "".join(pieces)


def text_writer(encoding: str = "utf-8") -> Writer:
    """One `str(item)` per line."""
    def writer(path: str, data: Iterable[Any]) -> None:
        with open(path, "w", encoding=encoding) as handle:
            handle.writelines(str(item) + "\n" for item in data)
        log("Result saved as plain text in '{}'".format(path))
    return writer


def tokenized_text_writer(encoding: str = "utf-8") -> Writer:
    """Token lists joined by single spaces."""
    plain = text_writer(encoding)
    return lambda path, data: plain(path, (" ".join(str(tok) for tok in sentence) for sentence in data))


def t2t_tokenized_text_writer(encoding: str = "utf-8") -> Writer:
    plain = text_writer(encoding)
    return lambda path, data: plain(path, t2t_detokenize(data))


# pylint: disable=invalid-name
UtfPlainTextWriter = tokenized_text_writer()
T2TWriter = t2t_tokenized_text_writer()


def __getattr__(name: str):
    # AutoWriter lived here before the writers were split like the reference's package
    if name in ("AutoWriter", "auto_writer"):
        from neuralmonkey_b200.writers import auto
        return getattr(auto, name)
    raise AttributeError(name)
