"""Exceptions of the INI configuration layer (reference: neuralmonkey/config/exceptions.py)."""
import traceback
from typing import Any, Optional


class ParseError(Exception):
    """Syntax error in an INI file; carries the line number when known."""

    def __init__(self, message: str, line: Optional[int] = None) -> None:
        super().__init__()
        self.message = message
        self.line = line

    def set_line(self, line: int) -> None:
        self.line = line

    def __str__(self) -> str:
        if self.line is not None:
            return "INI error on line {}: {}".format(self.line, self.message)
        return "INI parsing error: {}".format(self.message)


class ConfigInvalidValueException(Exception):
    def __init__(self, value: Any, message: str) -> None:
        super().__init__()
        self.value = value
        self.message = message

    def __str__(self) -> str:
        return "Error in configuration of {}: {}".format(self.value, self.message)


class ConfigBuildException(Exception):
    """An object of the configuration failed to build."""

    def __init__(self, object_name: Any, original_exception: Exception) -> None:
        super().__init__()
        self.object_name = object_name
        self.original_exception = original_exception

    def __str__(self) -> str:
        trc = "".join(traceback.format_list(traceback.extract_tb(
            self.original_exception.__traceback__)))
        return "Error while loading '{}': {}\nTraceback: {}".format(
            self.object_name, self.original_exception, trc)
