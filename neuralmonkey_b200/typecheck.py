"""Run-time check of a constructor's arguments against its annotations.

Every constructor of the reference starts with typeguard's `check_argument_types()`, so a configuration
that passes a string where a size is expected fails with a `TypeError` while the model is being built
(error convention).  typeguard is not a dependency here; this is the small part of it
those calls rely on: plain classes, `Optional` / `Union`, `List` / `Tuple` / `Dict` / `Set` with element
types, `Callable`, `Any`; `int` where `float` is expected is accepted (PEP 484), `bool` for a number is
not; a default of `None` makes the parameter optional; annotations that cannot be resolved are skipped."""
import collections.abc
import inspect
import sys
import typing
from typing import Any


def _matches(value: Any, expected: Any) -> bool:
    if expected is Any or expected is inspect.Parameter.empty or isinstance(expected, (str, typing.TypeVar)):
        return True
    origin, args = typing.get_origin(expected), typing.get_args(expected)
    if origin is typing.Union:
        return any(_matches(value, a) for a in args)
    if origin in (list, set, frozenset, collections.abc.Sequence, collections.abc.Iterable):
        container = {list: list, set: set, frozenset: frozenset}.get(origin, (list, tuple))
        return isinstance(value, container) and (not args or all(_matches(v, args[0]) for v in value))
    if origin is tuple:
        if not isinstance(value, tuple):
            return False
        if len(args) == 2 and args[1] is Ellipsis:
            return all(_matches(v, args[0]) for v in value)
        return not args or (len(value) == len(args) and all(_matches(v, a) for v, a in zip(value, args)))
    if origin is dict:
        return isinstance(value, dict) and (not args or all(
            _matches(k, args[0]) and _matches(v, args[1]) for k, v in value.items()))
    if origin is collections.abc.Callable or expected is typing.Callable:
        return callable(value)
    if origin is type:
        return isinstance(value, type)
    if origin is not None:                      # a generic this helper does not know: do not guess
        return True
    if expected is float:
        return isinstance(value, (int, float)) and not isinstance(value, bool)
    if expected is int:
        return isinstance(value, int) and not isinstance(value, bool)
    if isinstance(expected, type):
        return isinstance(value, expected)
    return True


def check_argument_types() -> bool:
    """Validate the arguments of the CALLING function or method; raises TypeError naming the argument."""
    frame = sys._getframe(1)
    code, local = frame.f_code, frame.f_locals
    func = None
    owner = local.get("self")
    if owner is not None:
        for klass in type(owner).__mro__:
            candidate = klass.__dict__.get(code.co_name)
            candidate = getattr(candidate, "__func__", candidate)
            if getattr(candidate, "__code__", None) is code:
                func = candidate
                break
    if func is None:
        candidate = frame.f_globals.get(code.co_name)
        if getattr(candidate, "__code__", None) is code:
            func = candidate
    if func is None:
        return True
    try:
        hints = typing.get_type_hints(func)
    except Exception:  # pylint: disable=broad-except
        hints = dict(getattr(func, "__annotations__", {}))
    parameters = inspect.signature(func).parameters
    for name, expected in hints.items():
        if name == "return" or name not in local or name not in parameters:
            continue
        value = local[name]
        if value is None and parameters[name].default is None:
            continue
        if not _matches(value, expected):
            raise TypeError('type of argument "{}" must be {}; got {} instead'.format(
                name, getattr(expected, "__name__", str(expected)), type(value).__name__))
    return True


def check_type(name: str, value: Any, expected: Any) -> bool:
    """One value against one annotation (typeguard's `check_type`), same message as for an argument."""
    if not _matches(value, expected):
        raise TypeError('type of argument "{}" must be {}; got {} instead'.format(
            name, getattr(expected, "__name__", str(expected)), type(value).__name__))
    return True
