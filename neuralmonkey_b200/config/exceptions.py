"""Errors raised while reading and building an experiment configuration
(same names and messages as neuralmonkey/config/exceptions.py, so user-facing output and
`except` clauses written against the reference keep working)."""
import traceback
from typing import Any, Optional


class _ConfigError(Exception):
    """Common plumbing: the message is rendered lazily by `describe()`."""

    def describe(self) -> str:
        raise NotImplementedError

    def __str__(self) -> str:
        return self.describe()


class ParseError(_ConfigError):
    """A syntax error in an INI file.  The parser attaches the line once it is known."""

    def __init__(self, message: str, line: Optional[int] = None) -> None:
        _ConfigError.__init__(self)
        self.message, self.line = message, line

    def set_line(self, line: int) -> None:
        self.line = line

    def describe(self) -> str:
        where = "INI parsing error" if self.line is None else "INI error on line {}".format(self.line)
        return "{}: {}".format(where, self.message)


class ConfigInvalidValueException(_ConfigError):
    """A value that parsed but cannot be used where it stands."""

    def __init__(self, value: Any, message: str) -> None:
        _ConfigError.__init__(self)
        self.value, self.message = value, message

    def describe(self) -> str:
        return "Error in configuration of {}: {}".format(self.value, self.message)


class ConfigBuildException(_ConfigError):
    """Wraps whatever a constructor raised while an INI section was being instantiated."""

    def __init__(self, object_name: Any, original_exception: Exception) -> None:
        _ConfigError.__init__(self)
        self.object_name, self.original_exception = object_name, original_exception

    def describe(self) -> str:
        frames = traceback.extract_tb(self.original_exception.__traceback__)
        return "Error while loading '{}': {}\nTraceback: {}".format(
            self.object_name, self.original_exception, "".join(traceback.format_list(frames)))
