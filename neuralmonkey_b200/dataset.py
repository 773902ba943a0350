"""Datasets and batching (behaviour of neuralmonkey/dataset.py:55-615).

A Dataset is a set of equally long named data series behind iterator factories; `batches()`
yields small in-memory Datasets, optionally bucketed by the longest series of each example
and optionally through a lazily refilled (and shuffled) buffer.
"""
import glob
import os
import random
from collections import deque
from itertools import islice
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple, Union

from neuralmonkey_b200.logging import debug, log, warn
from neuralmonkey_b200.readers.plain_text_reader import UtfPlainTextReader
from neuralmonkey_b200.writers.auto import AutoWriter
from neuralmonkey_b200.writers.plain_text_writer import Writer

Reader = Callable[[List[str]], Any]


class BatchingScheme:
    def __init__(self, batch_size: int = None, drop_remainder: bool = False,
                 bucket_boundaries: List[int] = None, bucket_batch_sizes: List[int] = None,
                 ignore_series: List[str] = None) -> None:
        """Either a fixed `batch_size` or length buckets (`bucket_boundaries` are upper limits,
        `bucket_batch_sizes` has one more entry for the overflow bucket)."""
        self.batch_size = batch_size
        self.drop_remainder = drop_remainder
        self.bucket_boundaries = bucket_boundaries
        self.bucket_batch_sizes = bucket_batch_sizes
        self.ignore_series = list(ignore_series) if ignore_series is not None else []
        if (self.batch_size is None) == (self.bucket_boundaries is None):
            raise ValueError("You must specify either batch_size or bucket_boundaries, not both")
        if self.bucket_boundaries is not None:
            if self.bucket_batch_sizes is None:
                raise ValueError("You must specify bucket_batch_sizes")
            if len(self.bucket_batch_sizes) != len(self.bucket_boundaries) + 1:
                raise ValueError("There should be N+1 batch sizes for N bucket boundaries")


def _expand_patterns(patterns: Union[str, List[str]]) -> List[str]:
    if isinstance(patterns, str):
        patterns = [patterns]
    paths = []  # type: List[str]
    for pattern in patterns:
        matched = sorted(glob.glob(pattern))
        if not matched:
            raise FileNotFoundError("Pattern did not match any files: {}".format(pattern))
        paths.extend(matched)
    return paths


def _is_file_spec(spec: Any) -> bool:
    """str | [str] | (str | [str], reader)."""
    def is_files(x):
        return isinstance(x, str) or (isinstance(x, list) and all(isinstance(i, str) for i in x))
    if is_files(spec):
        return True
    return isinstance(spec, tuple) and len(spec) == 2 and is_files(spec[0]) and callable(spec[1])


def load(name: str, series: List[str], data: List[Any], batching: BatchingScheme = None,
         outputs: List[Tuple] = None, buffer_size: int = None, shuffled: bool = False) -> "Dataset":
    """Create a dataset (dataset.py:207-326).  Each entry of `data` is a file spec, a
    series-level preprocessor `(function, source series)` or a dataset-level preprocessor
    `function(iterators)`."""
    if batching is None:
        from neuralmonkey_b200.experiment import Experiment
        log("Using default batching scheme for dataset {}.".format(name))
        batch_size = Experiment.get_current().config.args.batch_size
        if batch_size is None:
            raise ValueError("Argument main.batch_size is not specified, cannot use default "
                             "batching scheme.")
        batching = BatchingScheme(batch_size=batch_size)
    if not series:
        raise ValueError("No dataset series specified.")
    if not [s for s in data if _is_file_spec(s)]:
        raise ValueError("At least one data series should be from a file")
    if len(series) != len(data):
        raise ValueError("The 'series' and 'data' lists should have the same number of elements: "
                         "{} vs {}.".format(len(series), len(data)))
    if len(series) != len(set(series)):
        raise ValueError("There are duplicate series.")
    if outputs is not None and len({o[0] for o in outputs}) != len(outputs):
        raise ValueError("Multiple outputs for a single series")
    log("Initializing dataset {}.".format(name))
    iterators = {}  # type: Dict[str, Callable[[], Iterator]]
    series_level, dataset_level = {}, {}
    for s_name, spec in zip(series, data):
        if _is_file_spec(spec):
            files, reader = (spec[0], spec[1]) if isinstance(spec, tuple) else (spec, UtfPlainTextReader)
            files = _expand_patterns(files)
            for path in files:
                if not os.path.isfile(path):
                    raise FileNotFoundError("File not found. Series: {}, Path: {}".format(s_name, path))
            iterators[s_name] = (lambda r=reader, f=files: r(f))
        elif isinstance(spec, tuple) and len(spec) == 2 and callable(spec[0]) and isinstance(spec[1], str):
            series_level[s_name] = spec
        else:
            if not callable(spec):
                raise ValueError("Unrecognised data source for series '{}'".format(s_name))
            dataset_level[s_name] = spec
    for s_name, (prep, source) in series_level.items():
        if source not in iterators:
            raise ValueError("Source series for series-level preprocessor nonexistent: Preprocessed "
                             "series '{}', source series '{}'".format(s_name, source))
        iterators[s_name] = (lambda p=prep, s=source: (p(item) for item in iterators[s]()))
    for s_name, func in dataset_level.items():
        iterators[s_name] = (lambda f=func: f(iterators))
    output_dict = None
    if outputs is not None:
        output_dict = {}
        for out in outputs:
            output_dict[out[0]] = (out[1], out[2] if len(out) > 2 else AutoWriter)
    buf = (buffer_size // 2, buffer_size) if buffer_size is not None else None
    return Dataset(name, iterators, batching, output_dict, buf, shuffled)


def load_dataset_from_files(name: str = None, lazy: bool = False, preprocessors: List[Tuple] = None,
                            **kwargs) -> "Dataset":
    """Compat shim for the pre-0.3 API still used by examples/translation.ini:59-71:
    `s_<series>=path`, `s_<series>_out=path`, `preprocessors=[(source, new, function)]`."""
    series, data, outputs = [], [], []
    for key, value in kwargs.items():
        if key.startswith("s_") and key.endswith("_out"):
            outputs.append((key[2:-4], value))
        elif key.startswith("s_"):
            series.append(key[2:])
            data.append(value)
        elif key.startswith("pre_"):
            raise ValueError("series-level 'pre_' preprocessors are not supported by the shim")
    for src, new, func in preprocessors or []:
        series.append(new)
        data.append((func, src))
    return load(name or "dataset", series, data, outputs=outputs or None,
                buffer_size=(5000 if lazy else None))


class Dataset:
    def __init__(self, name: str, iterators: Dict[str, Callable[[], Iterator]],
                 batching: BatchingScheme, outputs: Dict[str, Tuple[str, Writer]] = None,
                 buffer_size: Tuple[int, int] = None, shuffled: bool = False) -> None:
        self.name = name
        self.iterators = iterators
        self.batching = batching
        self.outputs = outputs
        self.lazy = buffer_size is not None
        if self.lazy:
            self.buffer_min_size, self.buffer_size = buffer_size
        self.shuffled = shuffled
        self.length = None  # type: Optional[int]
        if not self.lazy:
            data = {s_name: list(it()) for s_name, it in self.iterators.items()}
            lengths = {s_name: len(s_data) for s_name, s_data in data.items()}
            if len(set(lengths.values())) > 1:
                raise ValueError("Lengths of data series do not match: {}".format(str(lengths)))
            self.length = next(iter(lengths.values())) if lengths else 0
            self.iterators = {s_name: (lambda n=s_name: iter(data[n])) for s_name in self.iterators}

    def __len__(self) -> int:
        if self.lazy:
            raise NotImplementedError("Querying the len of a lazy dataset.")
        return self.length

    def __contains__(self, name: str) -> bool:
        return name in self.iterators

    @property
    def series(self) -> List[str]:
        return list(sorted(self.iterators.keys()))

    def get_series(self, name: str) -> Iterator:
        return self.iterators[name]()

    def maybe_get_series(self, name: str) -> Optional[Iterator]:
        return self.get_series(name) if name in self.iterators else None

    def _bucket_of(self, row: Dict[str, Any]) -> int:
        if self.batching.bucket_boundaries is None:
            return 0
        length = max(len(row[key]) for key in row if key not in self.batching.ignore_series)
        best = -1
        for b_id, limit in enumerate(self.batching.bucket_boundaries):
            if length <= limit and (best == -1 or limit < self.batching.bucket_boundaries[best]):
                best = b_id
        return best  # -1 = overflow bucket (last list)

    def batches(self) -> Iterator["Dataset"]:
        """Yield batch datasets (dataset.py:467-560)."""
        scheme = self.batching
        max_bs = scheme.batch_size if scheme.batch_size is not None else max(scheme.bucket_batch_sizes)
        if self.lazy and self.buffer_min_size < max_bs:
            warn("Minimum buffer size ({}) lower than batch size ({}). It is recommended to use "
                 "large buffer size.".format(self.buffer_min_size, max_bs))
        its = {s: it() for s, it in self.iterators.items()}
        rows = (dict(zip(its, row)) for row in zip(*its.values()))
        lbuf = list(islice(rows, self.buffer_size)) if self.lazy else list(rows)
        if self.shuffled:
            random.shuffle(lbuf)
        buf = deque(lbuf)
        n_buckets = 1 + (len(scheme.bucket_boundaries) if scheme.bucket_boundaries is not None else 0)
        buckets = [[] for _ in range(n_buckets)]  # type: List[List[Dict[str, Any]]]
        batch_index = 0

        def make_batch(bucket_rows):
            nonlocal batch_index
            data = {key: (lambda r=bucket_rows, k=key: (row[k] for row in r)) for key in bucket_rows[0]}
            batch = Dataset("{}.batch.{}".format(self.name, batch_index), data, scheme)
            batch_index += 1
            return batch

        while buf:
            row = buf.popleft()
            b_id = self._bucket_of(row)
            buckets[b_id].append(row)
            limit = scheme.batch_size if scheme.bucket_batch_sizes is None else scheme.bucket_batch_sizes[b_id]
            if len(buckets[b_id]) >= limit:
                yield make_batch(buckets[b_id])
                buckets[b_id] = []
            if self.lazy and len(buf) < self.buffer_min_size:
                buf.extend(islice(rows, self.buffer_size - len(buf)))
                if self.shuffled:
                    tmp = list(buf)
                    random.shuffle(tmp)
                    buf = deque(tmp)
        if not scheme.drop_remainder:
            for bucket in buckets:
                if bucket:
                    yield make_batch(bucket)

    def subset(self, start: int, length: int) -> "Dataset":
        outputs = None
        if self.outputs is not None:
            outputs = {key: ("{}.{:010}".format(path, start), writer)
                       for key, (path, writer) in self.outputs.items()}
        slices = {s_id: (lambda s=s_id: islice(self.get_series(s), start, start + length))
                  for s_id in self.iterators}
        return Dataset("{}.{}.{}".format(self.name, start, length), slices, self.batching, outputs,
                       (self.buffer_min_size, self.buffer_size) if self.lazy else None, self.shuffled)
