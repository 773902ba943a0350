"""INI parser of the experiment configuration (grammar of neuralmonkey/config/parsing.py:16-258).

Value grammar, tried in this order on the stripped string:
    True / False / None
    integer            -?[0-9]+
    float              -?[0-9]*.[0-9]*(e[+-]?[0-9]+)?   or   -?[0-9]+e[+-]?[0-9]+
    "string"           with {var} templating against the [vars] section
    $var               value of a [vars] entry, falling back to the environment
    dotted.name        class / function symbol (resolved lazily by the builder)
    <object[.attr]>    reference to another section (optionally an attribute chain)
    [a, b, ...]        list (nested brackets allowed)
    (a, b, ...)        tuple
Every value remembers its line so errors can name it; `-s section.option=value` style
overrides are applied before values are parsed.  The builtin variable TIME holds the
start time.
"""
import configparser
import os
import re
import time
from collections import OrderedDict
from typing import IO, Any, Dict, Iterable, List, Optional, Tuple

from neuralmonkey_b200.config.builder import ClassSymbol, ObjectRef
from neuralmonkey_b200.config.exceptions import ParseError
from neuralmonkey_b200.logging import log

_INTEGER = re.compile(r"^-?[0-9]+$")
_FLOAT = re.compile(r"^-?[0-9]*\.[0-9]*(e[+-]?[0-9]+)?$|^-?[0-9]+e[+-]?[0-9]+$")
_STRING = re.compile(r'^"(.*)"$')
_VAR_REF = re.compile(r"^\$([a-zA-Z][a-zA-Z0-9_]*)$")
_OBJECT_REF = re.compile(r"^<([a-zA-Z][a-zA-Z0-9_]*(\.[a-zA-Z][a-zA-Z0-9_]*)*)>$")
_CLASS_NAME = re.compile(r"^_*[a-zA-Z][a-zA-Z0-9_]*(\._*[a-zA-Z][a-zA-Z0-9_]*)+$")
_LIST = re.compile(r"\[([^]]*)\]")
_TUPLE = re.compile(r"\(([^)]+)\)")
_LINE_SUFFIX = re.compile(r"^(.*) ([0-9]+)$")

_CONSTANTS = {"False": False, "True": True, "None": None}


class VarsDict(OrderedDict):
    """[vars] section; unknown names are looked up in the environment (parsed if possible)."""

    def __missing__(self, key: str) -> Any:
        if key in os.environ:
            raw = os.environ[key]
            try:
                value = parse_value(raw, self)
            except ParseError:
                value = raw
            log("Variable {}={!r} taken from the environment.".format(key, value))
            return value
        raise ParseError("Undefined variable: {}".format(key))


def split_on_commas(text: str) -> List[str]:
    """Split on top-level commas; commas inside () or [] are kept."""
    items = []  # type: List[str]
    current = []  # type: List[str]
    stack = []  # type: List[str]
    for col, char in enumerate(text):
        if char == "," and not stack:
            if current:
                items.append("".join(current))
            current = []
            continue
        if char == " " and not current:
            continue
        if char in "([":
            stack.append(char)
        elif char in ")]":
            opener = "(" if char == ")" else "["
            if not stack or stack.pop() != opener:
                raise ParseError("Invalid bracket end '{}', col {}.".format(char, col))
        current.append(char)
    if current:
        items.append("".join(current))
    return items


def parse_value(text: str, variables: VarsDict) -> Any:
    """Parse one value string according to the grammar in the module docstring."""
    text = text.strip()
    if text in _CONSTANTS:
        return _CONSTANTS[text]
    if _INTEGER.match(text):
        return int(text)
    if _FLOAT.match(text):
        return float(text)
    match = _STRING.match(text)
    if match:
        return match.group(1).format_map(variables)
    match = _VAR_REF.match(text)
    if match:
        return variables[match.group(1)]
    if _CLASS_NAME.match(text):
        return ClassSymbol(text)
    match = _OBJECT_REF.match(text)
    if match:
        return ObjectRef(match.group(1))
    match = _LIST.match(text)
    if match:
        # nested lists: the regex only finds the first ']' - take the whole bracketed text
        inner = text[1:text.rfind("]")] if text.startswith("[") else match.group(1)
        if not inner.strip():
            return []
        return [parse_value(item, variables) for item in split_on_commas(inner)]
    match = _TUPLE.match(text)
    if match:
        inner = text[1:text.rfind(")")] if text.startswith("(") else match.group(1)
        return tuple(parse_value(item, variables) for item in split_on_commas(inner))
    raise ParseError("Cannot parse value: '{}'.".format(text))


def _read_ini(lines: Iterable[str], filename: str = "") -> "OrderedDict[str, OrderedDict]":
    """configparser pass; every value is returned as (line number, raw string)."""
    numbered = (line.strip() + " " + str(i + 1) if line.strip() else ""
                for i, line in enumerate(lines))
    parser = configparser.ConfigParser()
    parser.read_file(numbered, source=filename)
    sections = OrderedDict()  # type: OrderedDict
    for section in parser.sections():
        sections[section] = OrderedDict()
        for key in parser[section]:
            match = _LINE_SUFFIX.match(parser[section][key])
            assert match is not None
            sections[section][key] = (match.group(2), match.group(1))
    return sections


def apply_change(config: Dict[str, Any], setting: str) -> None:
    """`section.option=value` (or `option=value` for [main]) command-line override."""
    if "=" not in setting:
        raise ParseError("Invalid setting '{}'".format(setting))
    key, value = (s.strip() for s in setting.split("=", maxsplit=1))
    section, option = key.split(".", maxsplit=1) if "." in key else ("main", key)
    if section not in config:
        log("Creating new section '{}'".format(section))
        config[section] = OrderedDict()
    config[section][option] = (-1, value)


def parse_file(config_file: Iterable[str],
               changes: Optional[Iterable[str]] = None) -> Tuple[Dict[str, Any], Dict[str, Any]]:
    """Returns (raw strings per section, parsed values per section)."""
    config = _read_ini(config_file)
    for change in changes or []:
        apply_change(config, change)
    variables = VarsDict()
    variables["TIME"] = time.strftime("%Y-%m-%d-%H-%M-%S")

    def parse_section(name: str, out: Dict[str, Any]) -> None:
        for key, (lineno, raw) in config[name].items():
            try:
                out[key] = parse_value(raw, variables)
            except ParseError as exc:
                exc.set_line(lineno)
                raise

    if "vars" in config:
        parse_section("vars", variables)
    parsed = OrderedDict()  # type: Dict[str, Any]
    for name in config:
        if name != "vars":
            parsed[name] = OrderedDict()
            parse_section(name, parsed[name])
    raw = OrderedDict((name, OrderedDict((k, v) for k, (_, v) in sec.items()))
                      for name, sec in config.items())
    return raw, parsed


def write_file(config_dict: Dict[str, Any], config_file: IO[str]) -> None:
    parser = configparser.ConfigParser()
    parser.read_dict(config_dict)
    parser.write(config_file, space_around_delimiters=False)


# the reference's (private) names for the same functions: neuralmonkey/tests/test_config.py uses them
_parse_value = parse_value
_split_on_commas = split_on_commas
