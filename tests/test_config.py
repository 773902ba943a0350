"""The INI layer (the drop-in boundary) against fixtures produced by the reference's own
parser (tests/golden/make_config_golden.py), plus builder behaviour."""
import json
import os
import re

import pytest

from neuralmonkey_b200.config import parsing
from neuralmonkey_b200.config.builder import ClassSymbol, ObjectRef, build_config
from neuralmonkey_b200.config.exceptions import ConfigBuildException, ParseError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_golden.json")


def encode(value):
    if isinstance(value, ClassSymbol):
        return {"__class__": value.clazz}
    if isinstance(value, ObjectRef):
        return {"__ref__": value.expression}
    if isinstance(value, tuple):
        return {"__tuple__": [encode(v) for v in value]}
    if isinstance(value, list):
        return [encode(v) for v in value]
    if isinstance(value, dict):
        return {k: encode(v) for k, v in value.items()}
    return value


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN, encoding="utf-8"))


def test_reference_inis_parse_identically(golden, monkeypatch):
    monkeypatch.setenv("NM_EXPERIMENT_NAME", "small")  # environment fallback of tests/small.ini
    assert len(golden["inis"]) >= 5
    for name, entry in golden["inis"].items():
        _raw, parsed = parsing.parse_file(entry["text"].splitlines(keepends=True))
        # the builtin TIME variable (tests/small.ini uses it) differs by construction
        stamp = re.compile(r"\d{4}(-\d{2}){5}")
        got = json.loads(stamp.sub("TIME", json.dumps(encode(parsed), sort_keys=True)))
        want = json.loads(stamp.sub("TIME", json.dumps(entry["parsed"], sort_keys=True)))
        assert got == want, name


def test_value_grammar_matches_reference(golden):
    variables = parsing.VarsDict()
    variables["TIME"] = "T0"
    for text, want in golden["values"].items():
        got = json.loads(json.dumps(encode(parsing.parse_value(text, variables))))
        assert got == want, text


def test_parse_errors_carry_line_numbers():
    ini = ["[main]\n", "a=1\n", "b=@@\n"]
    with pytest.raises(ParseError) as err:
        parsing.parse_file(ini)
    assert "line 3" in str(err.value)
    with pytest.raises(ParseError):
        parsing.parse_value("$undefined_variable_xyz", parsing.VarsDict())
    with pytest.raises(ParseError):
        parsing.split_on_commas("(1, 2]")


def test_changes_vars_and_env(monkeypatch):
    monkeypatch.setenv("NMB_TEST_SIZE", "17")
    ini = ["[vars]\n", "dim=8\n", "[main]\n", "x=$dim\n", "y=\"out-{dim}\"\n", "z=$NMB_TEST_SIZE\n"]
    _raw, parsed = parsing.parse_file(ini, changes=["main.x=9", "other.k=[1, 2]"])
    assert parsed["main"] == {"x": 9, "y": "out-8", "z": 17}
    assert parsed["other"]["k"] == [1, 2]


def test_builder_instantiates_shares_and_names_objects():
    ini = """
[main]
a=<first>
b=<second>
c=<second.attentions>
[first]
class=attention.Attention
encoder=<enc>
[second]
class=decoders.Decoder
encoders=[<enc>]
attentions=[<first>]
vocabulary=<vocab>
data_id="target"
max_output_len=5
rnn_size=8
embedding_size=8
[enc]
class=encoders.SentenceEncoder
vocabulary=<vocab>
data_id="source"
embedding_size=4
rnn_size=3
[vocab]
class=vocabulary.Vocabulary
words=["x", "y"]
[unused_section]
class=vocabulary.Vocabulary
words=[]
""".strip().splitlines(keepends=True)
    _raw, parsed = parsing.parse_file(ini)
    model, objects = build_config(parsed, set())
    assert model["a"].name == "first"                      # section name is the default `name`
    assert model["b"].attentions[0] is model["a"]          # references are shared, not copied
    assert model["c"] == [model["a"]]                      # attribute chains resolve
    assert objects["enc"].input_sequence.name == "enc_input"
    assert "unused_section" not in objects


def test_builder_reports_bad_arguments():
    ini = "[main]\nv=<vocab>\n[vocab]\nclass=vocabulary.Vocabulary\nnot_an_argument=3\n".splitlines(
        keepends=True)
    _raw, parsed = parsing.parse_file(ini)
    with pytest.raises(ConfigBuildException):
        build_config(parsed, set())


def test_tf_names_used_by_reference_configs_resolve():
    for name in ("tf.contrib.opt.LazyAdamOptimizer", "tf.train.AdamOptimizer",
                 "tf.random_uniform_initializer", "tf.tanh"):
        assert ClassSymbol(name).create() is not None


def test_builder_matches_reference_on_toy_plugin():
    """Object graphs built by the reference's own config/builder.py (golden/make_builder_golden.py)
    from INIs over tests/golden/toy_plugin/nmtoy.py: default names, shared references, attribute
    chains, tuples, callables, unused sections."""
    import json
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden", "toy_plugin"))
    try:
        import nmtoy
        golden = json.load(open(os.path.join(here, "golden", "builder_golden.json")))
        for name, text in golden["inis"].items():
            _raw, parsed = parsing.parse_file(text.strip().splitlines(keepends=True))
            model, objects = build_config(parsed, set())
            seen = {}
            got = {k: nmtoy.describe(v, seen) for k, v in sorted(model.items())}
            assert got == golden[name]["model"], name
            assert sorted(objects) == golden[name]["objects"], name
    finally:
        sys.path.pop(0)
