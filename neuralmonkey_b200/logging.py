"""Console + experiment.log logging (reference: neuralmonkey/logging.py:12-113).

Same entry points (`log`, `warn`, `notice`, `debug`, `log_print`); colours are plain ANSI
codes (the reference depends on `termcolor`).  `warn` raises under NEURALMONKEY_STRICT
(logging.py:59-65)."""
import os
import sys
import time
from typing import Any, Optional

_COLORS = {"red": "31", "green": "32", "yellow": "33", "blue": "34", "magenta": "35",
           "cyan": "36", "white": "37"}


def _colored(text: str, color: Optional[str]) -> str:
    if color is None or not sys.stderr.isatty():
        return text
    return "\033[{}m{}\033[0m".format(_COLORS.get(color, "0"), text)


class Logging:
    log_file = None
    strict_mode = bool(os.environ.get("NEURALMONKEY_STRICT"))
    debug_enabled_for = [s for s in os.environ.get("NEURALMONKEY_DEBUG_ENABLE", "").split(",") if s]
    debug_disabled_for = [s for s in os.environ.get("NEURALMONKEY_DEBUG_DISABLE", "").split(",") if s]
    quiet = bool(os.environ.get("NEURALMONKEY_QUIET"))

    @staticmethod
    def set_log_file(path: str) -> None:
        if Logging.log_file is not None:
            Logging.log_file.close()
        Logging.log_file = open(path, "w", encoding="utf-8", buffering=1)

    @staticmethod
    def log_print(text: str) -> None:
        if Logging.log_file is not None and not Logging.log_file.closed:
            Logging.log_file.write(text + "\n")
        if not Logging.quiet:
            print(text, file=sys.stderr)

    @staticmethod
    def log(message: str, color: str = "yellow") -> None:
        stamp = time.strftime("%Y-%m-%d %H:%M:%S")
        Logging.log_print("{}: {}".format(_colored(stamp, color), message))

    @staticmethod
    def notice(message: str) -> None:
        Logging.log("NOTICE: {}".format(message), color="red")

    @staticmethod
    def warn(message: str) -> None:
        if Logging.strict_mode:
            raise Exception("Encountered a warning in strict mode: " + message)
        Logging.log("WARNING: {}".format(message), color="red")

    @staticmethod
    def debug(message: str, label: Optional[str] = None) -> None:
        if not Logging.debug_enabled(label):
            return
        prefix = "DEBUG ({})".format(label) if label else "DEBUG"
        Logging.log("{}: {}".format(prefix, message), color="cyan")

    @staticmethod
    def debug_enabled(label: Optional[str] = None) -> bool:
        if label is None:
            label = "none"
        if "none" in Logging.debug_disabled_for and label == "none":
            return False
        if label in Logging.debug_disabled_for:
            return False
        return "all" in Logging.debug_enabled_for or label in Logging.debug_enabled_for


log = Logging.log
log_print = Logging.log_print
warn = Logging.warn
notice = Logging.notice
debug = Logging.debug
debug_enabled = Logging.debug_enabled
