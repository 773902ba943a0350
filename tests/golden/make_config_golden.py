"""Generate tests/golden/config_golden.json by running the REFERENCE's own INI parser
(/root/reference/neuralmonkey/config/parsing.py) on its own experiment configs and on a
list of value strings.  Run in the build container only (the reference is not on the GPU
box); the JSON it writes is the committed fixture the parity test reads.

The reference imports `termcolor` (absent here) through neuralmonkey.logging: a stub module
is injected; nothing else of the reference is patched.
"""
import json
import os
import sys
import types

REF = "/root/reference"
INIS = ["tests/bahdanau.ini", "examples/translation.ini", "tests/transformer.ini",
        "tests/beamsearch.ini", "tests/captioning.ini", "tests/small.ini"]
VALUES = ["42", "-7", "1.0e-8", "-.5", "3e4", "True", "None", '"plain"', '"a {TIME} b"',
          "encoders.recurrent.SentenceEncoder", "tf.contrib.opt.LazyAdamOptimizer",
          "<decoder>", "<decoder.vocabulary>", "[1, 2, 3]", "[]", '[("a", <x>), ("b", c.d)]',
          "(1, 2.5, \"s\")", '[<trainer1>, <trainer1>, <trainer2>]', "_private.Name", "1e3"]


def encode(value):
    from neuralmonkey.config.builder import ClassSymbol, ObjectRef
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


def main():
    stub = types.ModuleType("termcolor")
    stub.colored = lambda text, *a, **k: text
    sys.modules["termcolor"] = stub
    sys.path.insert(0, REF)
    os.environ["NEURALMONKEY_QUIET"] = "1"
    os.environ["NM_EXPERIMENT_NAME"] = "small"  # tests/small.ini reads it (tests_run.sh:32-38)
    from neuralmonkey.config import parsing
    out = {"inis": {}, "values": {}}
    for ini in INIS:
        text = open(os.path.join(REF, ini), encoding="utf-8").read()
        _raw, parsed = parsing.parse_file(text.splitlines(keepends=True))
        out["inis"][ini] = {"text": text, "parsed": encode(parsed)}
    vars_dict = parsing.VarsDict()
    vars_dict["TIME"] = "T0"
    for val in VALUES:
        # pylint: disable=protected-access
        out["values"][val] = encode(parsing._parse_value(val, vars_dict))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_golden.json")
    with open(dst, "w", encoding="utf-8") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()
