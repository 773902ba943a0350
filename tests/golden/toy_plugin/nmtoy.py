"""Toy plugin classes both the reference's and this package's INI builders can instantiate
(`class=nmtoy.Leaf`): used to pin the builder's semantics - section name as default `name`,
shared references, nested lists / tuples, attribute chains, callables, default arguments."""


class Leaf:
    def __init__(self, name: str, value: int = 7, tags: list = None) -> None:
        self.name, self.value, self.tags = name, value, tags


class Node:
    def __init__(self, name: str, children: list, pair: tuple = None, factory=None, scale: float = 1.0) -> None:
        self.name, self.children, self.pair, self.factory, self.scale = name, children, pair, factory, scale
        self.first = children[0] if children else None


def make_leaf(value: int = 1) -> Leaf:
    return Leaf("made", value)


def describe(obj, seen=None):
    """A JSON-able picture of an object graph with identities (shared objects get the same id)."""
    seen = {} if seen is None else seen
    if isinstance(obj, (Leaf, Node)):
        if id(obj) in seen:
            return {"ref": seen[id(obj)]}
        seen[id(obj)] = len(seen)
        fields = {k: describe(v, seen) for k, v in sorted(vars(obj).items())}
        return {"type": type(obj).__name__, "id": seen[id(obj)], "fields": fields}
    if isinstance(obj, (list, tuple)):
        return {"seq": type(obj).__name__, "items": [describe(v, seen) for v in obj]}
    if callable(obj):
        return {"callable": getattr(obj, "__name__", str(obj))}
    return obj
