"""argparse-like loader for the [main] section
(behaviour of neuralmonkey/config/configuration.py:10-120)."""
import traceback
from argparse import Namespace
from collections import OrderedDict
from typing import Any, Callable, List, Optional

from neuralmonkey_b200.config.builder import build_config
from neuralmonkey_b200.config.parsing import parse_file, write_file
from neuralmonkey_b200.logging import log


class Configuration:
    def __init__(self) -> None:
        self.names = []  # type: List[str]
        self.defaults = {}
        self.conditions = {}
        self.ignored = set()
        self.raw_config = OrderedDict()
        self.config_dict = OrderedDict()
        self.objects = None
        self.args = None
        self.model = None

    def add_argument(self, name: str, required: bool = False, default: Any = None,
                     cond: Callable[[Any], bool] = None) -> None:
        if name in self.names:
            raise Exception("Data filed defined multiple times.")
        self.names.append(name)
        if not required:
            self.defaults[name] = default
        if cond is not None:
            self.conditions[name] = cond

    def ignore_argument(self, name: str) -> None:
        self.ignored.add(name)

    def make_namespace(self, d_obj) -> Namespace:
        n_space = Namespace()
        for name, value in d_obj.items():
            if name in self.conditions and not self.conditions[name](value):
                code = self.conditions[name].__code__
                raise Exception("Value of field '{}' does not satisfy condition defined at {}:{}."
                                .format(name, code.co_filename, code.co_firstlineno))
            setattr(n_space, name, value)
        for name, value in self.defaults.items():
            if name not in n_space.__dict__:
                n_space.__dict__[name] = value
        return n_space

    def load_file(self, path: str, changes: Optional[List[str]] = None) -> None:
        log("Loading INI file: '{}'".format(path), color="blue")
        try:
            with open(path, "r", encoding="utf-8") as file:
                raw_config, config_dict = parse_file(file, changes)
            log("INI file is parsed.")
            self.raw_config.update(raw_config)
            self.config_dict.update(config_dict)
        except Exception as exc:  # pylint: disable=broad-except
            log("Failed to load INI file: {}".format(exc), color="red")
            traceback.print_exc()
            exit(1)
        if "main" in self.config_dict:
            self.args = self.make_namespace(self.config_dict["main"])

    def build_model(self, warn_unused: bool = False) -> None:
        log("Building model based on the config.")
        self._check_loaded_conf()
        try:
            model, self.objects = build_config(self.config_dict, self.ignored, warn_unused)
        except Exception as exc:  # pylint: disable=broad-except
            log("Failed to build model: {}".format(exc), color="red")
            traceback.print_exc()
            exit(1)
        log("Model built.")
        self.model = self.make_namespace(model)

    def _check_loaded_conf(self) -> None:
        missing = [n for n in self.names if n not in self.args.__dict__]
        if missing:
            raise Exception("Missing mandatory fields: {}".format(", ".join(missing)))
        unexpected = [n for n in self.config_dict["main"]
                      if n not in self.names and n not in self.ignored]
        if unexpected:
            raise Exception("Unexpected fields: {}".format(", ".join(unexpected)))

    def save_file(self, path: str) -> None:
        with open(path, "w", encoding="utf-8") as file:
            write_file(self.raw_config, file)
