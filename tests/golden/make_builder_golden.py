"""Object graphs built by the reference's own `config/builder.py` from toy INIs
(tests/golden/toy_plugin/nmtoy.py holds the classes).  python tests/golden/make_builder_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "toy_plugin"))
from make_host_golden import install_stubs  # noqa: E402

INIS = {
    "shared": """
[main]
root=<tree>
leaf=<shared_leaf>
first_of_tree=<tree.first>
value=<shared_leaf.value>
[tree]
class=nmtoy.Node
children=[<shared_leaf>, <other_leaf>, <shared_leaf>]
pair=(<other_leaf>, 3)
factory=nmtoy.make_leaf
scale=2.5e-1
[shared_leaf]
class=nmtoy.Leaf
value=11
tags=["a", "b"]
[other_leaf]
class=nmtoy.Leaf
name="explicit name"
[never_used]
class=nmtoy.Leaf
""",
    "nested": """
[main]
top=<outer>
numbers=[1, 2.5, "three", None, True]
[outer]
class=nmtoy.Node
children=[<inner>, <inner>]
pair=("x", <inner.first>)
[inner]
class=nmtoy.Node
children=[<leaf>]
[leaf]
class=nmtoy.Leaf
""",
}


def main():
    install_stubs()
    import nmtoy
    from neuralmonkey.config.builder import build_config
    from neuralmonkey.config.parsing import parse_file
    out = {"inis": INIS}
    for name, text in INIS.items():
        _raw, parsed = parse_file(text.strip().splitlines(keepends=True))
        model, objects = build_config(parsed, set())
        seen = {}
        out[name] = {"model": {k: nmtoy.describe(v, seen) for k, v in sorted(model.items())},
                     "objects": sorted(objects)}
    json.dump(out, open(os.path.join(HERE, "builder_golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out["shared"])[:600])


if __name__ == "__main__":
    main()
