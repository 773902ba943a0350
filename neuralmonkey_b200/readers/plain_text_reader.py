"""Plain-text readers (names and behaviour of neuralmonkey/readers/plain_text_reader.py).
A reader maps a list of file paths to an iterator over examples; for text an example is the list
of tokens of one line.  `.gz` files are read transparently."""
import csv
import gzip
import io
import sys
import unicodedata
from typing import Callable, Iterable, Iterator, List

from neuralmonkey_b200.logging import warn

PlainTextFileReader = Callable[[List[str]], Iterable[List[str]]]

csv.field_size_limit(sys.maxsize)


def _is_alnum(char: str) -> bool:
    """Unicode letters and numbers (categories L* and N*): the reference's ALNUM_CHARSET."""
    return unicodedata.category(char)[0] in "LN"


def _lines(path: str, encoding: str) -> Iterator[str]:
    if path.endswith(".gz"):
        with gzip.open(path, "r") as handle:
            for raw in handle:
                yield str(raw, "utf-8")
    else:
        with open(path, encoding=encoding) as handle:
            yield from handle


def string_reader(encoding: str = "utf-8") -> Callable[[List[str]], Iterable[str]]:
    """Raw lines (with their line ends) of all files."""
    def reader(files: List[str]) -> Iterable[str]:
        for path in files:
            yield from _lines(path, encoding)
    return reader


def tokenized_text_reader(encoding: str = "utf-8") -> PlainTextFileReader:
    """Whitespace-separated tokens of every line."""
    def reader(files: List[str]) -> Iterable[List[str]]:
        for line in string_reader(encoding)(files):
            yield line.strip().split()
    return reader


def t2t_tokenized_text_reader(encoding: str = "utf-8") -> PlainTextFileReader:
    """tensor2tensor-style tokens: maximal runs of alphanumeric / non-alphanumeric characters; a
    run that is exactly one space is dropped unless it opens the line; the last run is always kept."""
    def reader(files: List[str]) -> Iterable[List[str]]:
        for line in string_reader(encoding)(files):
            text = line.strip()
            runs, start = [], 0
            for pos in range(1, len(text)):
                if _is_alnum(text[pos]) != _is_alnum(text[pos - 1]):
                    runs.append((start, text[start:pos]))
                    start = pos
            tokens = [run for begin, run in runs if run != " " or begin == 0]
            tokens.append(text[start:])
            yield tokens
    return reader


def column_separated_reader(column: int, delimiter: str = "\t", quotechar: str = None,
                            encoding: str = "utf-8") -> PlainTextFileReader:
    """Tokens of one (1-based) column; every line is parsed on its own, blanks after a delimiter
    are skipped, a missing column gives an empty example (with a warning)."""
    def reader(files: List[str]) -> Iterable[List[str]]:
        expected = None
        for line in string_reader(encoding)(files):
            kwargs = ({"quoting": csv.QUOTE_NONE} if quotechar is None else {"quotechar": quotechar})
            rows = list(csv.reader(io.StringIO(line.strip()), delimiter=delimiter, skipinitialspace=True,
                                   **kwargs))
            fields = rows[0] if rows else []
            if expected is None:
                expected = len(fields)
            elif expected != len(fields):
                warn("A mismatch in number of columns. Expected {} got {}".format(expected, len(fields)))
            if len(fields) < column:
                warn("There is a missing column number {} in the dataset.".format(column))
                yield []
            else:
                yield fields[column - 1].split()
    return reader


def csv_reader(column: int):
    return column_separated_reader(column, delimiter=",", quotechar='"')


def tsv_reader(column: int):
    return column_separated_reader(column, delimiter="\t", quotechar=None)


def get_plain_text_reader(encoding: str = "utf-8") -> PlainTextFileReader:
    """Older name of `tokenized_text_reader`."""
    return tokenized_text_reader(encoding)


# pylint: disable=invalid-name
UtfPlainTextReader = tokenized_text_reader()
T2TReader = t2t_tokenized_text_reader()
