"""Plain-text readers (reference: neuralmonkey/readers/plain_text_reader.py:10-120).
A reader maps a list of file paths to an iterator over examples; for text the example is
the list of whitespace-separated tokens of a line."""
import csv
import gzip
import io
import sys
from typing import Callable, Iterable, List

PlainTextFileReader = Callable[[List[str]], Iterable[List[str]]]


def _open(path: str, encoding: str):
    if path.endswith(".gz"):
        return io.TextIOWrapper(gzip.open(path, "r"), encoding=encoding)
    return open(path, encoding=encoding)


def get_plain_text_reader(encoding: str = "utf-8") -> PlainTextFileReader:
    def reader(files: List[str]) -> Iterable[List[str]]:
        for path in files:
            with _open(path, encoding) as f_data:
                for line in f_data:
                    yield line.strip().split()
    return reader


def column_separated_reader(column: int, delimiter: str = "\t", quotechar: str = None,
                            encoding: str = "utf-8") -> PlainTextFileReader:
    """Tokens of one (1-based) column of delimiter-separated files."""
    def reader(files: List[str]) -> Iterable[List[str]]:
        csv.field_size_limit(sys.maxsize)
        for path in files:
            with _open(path, encoding) as f_data:
                for row in csv.reader(f_data, delimiter=delimiter, quotechar=quotechar,
                                      quoting=csv.QUOTE_NONE if quotechar is None else csv.QUOTE_MINIMAL):
                    yield row[column - 1].strip().split() if len(row) >= column else []
    return reader


def csv_reader(column: int):
    return column_separated_reader(column=column, delimiter=",", quotechar='"')


def tsv_reader(column: int):
    return column_separated_reader(column=column, delimiter="\t", quotechar=None)


# pylint: disable=invalid-name
UtfPlainTextReader = get_plain_text_reader()
T2TReader = UtfPlainTextReader
