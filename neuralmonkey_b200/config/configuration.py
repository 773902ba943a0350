"""Declaration and loading of the `[main]` section of an experiment
(behaviour of neuralmonkey/config/configuration.py: declared fields with defaults and
conditions, unknown or missing fields are errors, failures are logged and end the process)."""
import sys
import traceback
from argparse import Namespace
from collections import OrderedDict
from typing import Any, Callable, Dict, List, NamedTuple, Optional

from neuralmonkey_b200.config.builder import build_config
from neuralmonkey_b200.config.parsing import parse_file, write_file
from neuralmonkey_b200.logging import log

_Field = NamedTuple("_Field", [("required", bool), ("default", Any),
                               ("condition", Optional[Callable[[Any], bool]])])


def _fail(stage: str, exc: Exception) -> None:
    log("Failed to {}: {}".format(stage, exc), color="red")
    traceback.print_exc()
    sys.exit(1)


class Configuration:
    def __init__(self) -> None:
        self._fields = OrderedDict()  # type: Dict[str, _Field]
        self.ignored = set()
        self.raw_config = OrderedDict()
        self.config_dict = OrderedDict()
        self.objects = None
        self.args = None
        self.model = None

    # -- declaration ---------------------------------------------------------------------------
    @property
    def names(self) -> List[str]:
        return list(self._fields)

    @property
    def defaults(self) -> Dict[str, Any]:
        return {n: f.default for n, f in self._fields.items() if not f.required}

    @property
    def conditions(self) -> Dict[str, Callable[[Any], bool]]:
        return {n: f.condition for n, f in self._fields.items() if f.condition is not None}

    def add_argument(self, name: str, required: bool = False, default: Any = None,
                     cond: Callable[[Any], bool] = None) -> None:
        if name in self._fields:
            raise Exception("Data filed defined multiple times.")
        self._fields[name] = _Field(required, default, cond)

    def ignore_argument(self, name: str) -> None:
        self.ignored.add(name)

    # -- values -----------------------------------------------------------------------------------
    def make_namespace(self, d_obj: Dict[str, Any]) -> Namespace:
        values = dict(self.defaults)
        for name, value in d_obj.items():
            field = self._fields.get(name)
            if field is not None and field.condition is not None and not field.condition(value):
                code = field.condition.__code__
                raise Exception("Value of field '{}' does not satisfy condition defined at {}:{}."
                                .format(name, code.co_filename, code.co_firstlineno))
            values[name] = value
        return Namespace(**values)

    def load_file(self, path: str, changes: Optional[List[str]] = None) -> None:
        log("Loading INI file: '{}'".format(path), color="blue")
        try:
            with open(path, "r", encoding="utf-8") as handle:
                raw, parsed = parse_file(handle, changes)
        except Exception as exc:  # pylint: disable=broad-except
            _fail("load INI file", exc)
        log("INI file is parsed.")
        self.raw_config.update(raw)
        self.config_dict.update(parsed)
        if "main" in self.config_dict:
            self.args = self.make_namespace(self.config_dict["main"])

    def build_model(self, warn_unused: bool = False) -> None:
        log("Building model based on the config.")
        self._check_loaded_conf()
        try:
            built, self.objects = build_config(self.config_dict, self.ignored, warn_unused)
        except Exception as exc:  # pylint: disable=broad-except
            _fail("build model", exc)
        log("Model built.")
        self.model = self.make_namespace(built)

    def _check_loaded_conf(self) -> None:
        given = set(vars(self.args))
        missing = [name for name in self._fields if name not in given]
        if missing:
            raise Exception("Missing mandatory fields: {}".format(", ".join(missing)))
        unexpected = [name for name in self.config_dict["main"]
                      if name not in self._fields and name not in self.ignored]
        if unexpected:
            raise Exception("Unexpected fields: {}".format(", ".join(unexpected)))

    def save_file(self, path: str) -> None:
        with open(path, "w", encoding="utf-8") as handle:
            write_file(self.raw_config, handle)
