"""The reference's OWN unit tests (neuralmonkey/tests/test_*.py of /root/reference), run unchanged against
this package: an import alias maps `neuralmonkey[.x]` to `neuralmonkey_b200[.x]`, the test files are
loaded from the reference tree and their `unittest` cases executed.  Covered are the files whose subject
is on the hot path or its host side and that need no TensorFlow session: constructor validation of the
decoder, BLEU, chrF, the dataset, the INI value parser.  (Running them is what exposed that the product's
BLEU lacked the reference's smoothing.)

Skipped when /root/reference is not mounted (build container only)."""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types
import unittest

import pytest

REFERENCE_TESTS = "/root/reference/neuralmonkey/tests"
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason="the reference tree is not mounted")


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, module):
        self.module = module

    def create_module(self, spec):
        return self.module

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    """`import neuralmonkey.x.y` hands out the very module object `neuralmonkey_b200.x.y`."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith("neuralmonkey.tests") or not (fullname == "neuralmonkey" or fullname.startswith("neuralmonkey.")):
            return None
        try:
            module = importlib.import_module("neuralmonkey_b200" + fullname[len("neuralmonkey"):])
        except ImportError:
            return None
        return importlib.machinery.ModuleSpec(fullname, _AliasLoader(module), is_package=hasattr(module, "__path__"))


@pytest.fixture
def reference_tests(monkeypatch):
    finder = _AliasFinder()
    sys.meta_path.insert(0, finder)
    saved = {k: v for k, v in sys.modules.items() if k == "tensorflow" or k == "neuralmonkey" or k.startswith("neuralmonkey.")}
    for name in saved:
        del sys.modules[name]
    tf = types.ModuleType("tensorflow")              # the tests call tf.reset_default_graph() around cases
    tf.reset_default_graph = lambda: None
    tf.test = types.SimpleNamespace(TestCase=unittest.TestCase)

    class Graph:                                     # test_vocabulary builds its vocabulary inside one
        def as_default(self):
            import contextlib
            return contextlib.nullcontext()
    tf.Graph = Graph
    sys.modules["tensorflow"] = tf
    package = types.ModuleType("neuralmonkey.tests")
    package.__path__ = [REFERENCE_TESTS]
    sys.modules["neuralmonkey.tests"] = package
    monkeypatch.setenv("NMB200_UNVERIFIED", "1")     # test_decoder constructs an LSTM decoder
    monkeypatch.chdir("/root/reference")             # test_dataset names tests/data/... relative to the tree

    def run(name, only=None, stubs=()):
        for stub_name, attrs in stubs:            # modules of the reference that are outside this package
            stub = types.ModuleType(stub_name)
            stub.__dict__.update(attrs)
            sys.modules[stub_name] = stub
        spec = importlib.util.spec_from_file_location("neuralmonkey.tests." + name,
                                                      os.path.join(REFERENCE_TESTS, name + ".py"))
        module = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = module
        spec.loader.exec_module(module)
        result = unittest.TestResult()
        loader = unittest.defaultTestLoader
        suite = loader.loadTestsFromNames(only, module) if only else loader.loadTestsFromModule(module)
        suite.run(result)
        return result
    yield run
    sys.meta_path.remove(finder)
    for name in [k for k in sys.modules if k == "tensorflow" or k == "neuralmonkey" or k.startswith("neuralmonkey.")]:
        del sys.modules[name]
    sys.modules.update(saved)


@pytest.mark.parametrize("name,cases", [("test_decoder", 5), ("test_bleu", 5), ("test_chrf", 6), ("test_dataset", 10),
                                        ("test_config", 4)])
def test_reference_unit_test_file_passes(reference_tests, name, cases):
    result = reference_tests(name)
    problems = ["{}: {}".format(case.id().split(".")[-1], trace.strip().splitlines()[-1])
                for case, trace in result.failures + result.errors]
    assert not problems, "\n".join(problems)
    assert result.testsRun >= cases


def test_reference_sentence_encoder_constructor_test_passes(reference_tests):
    """neuralmonkey/tests/test_encoders_init.py::test_sentence_encoder - every good / bad constructor
    argument combination of SentenceEncoder (the file's other case is about the sentence CNN encoder, which
    is outside the hot path: its module is stubbed so that the file imports)."""
    result = reference_tests("test_encoders_init", only=["TestEncodersInit.test_sentence_encoder"],
                             stubs=[("neuralmonkey.encoders.sentence_cnn_encoder", {"SentenceCNNEncoder": object})])
    problems = [trace.strip().splitlines()[-1] for _case, trace in result.failures + result.errors]
    assert not problems and result.testsRun == 1, problems


def test_reference_vocabulary_tests_pass(reference_tests):
    """neuralmonkey/tests/test_vocabulary.py without its two session-bound cases (string -> index lookup
    tables are TensorFlow ops there; here the conversion is host code, covered by the golden tests)."""
    result = reference_tests("test_vocabulary", only=["TestVocabulary.test_all_words_in", "TestVocabulary.test_unknown_word",
                                                      "TestVocabulary.test_padding", "TestVocabulary.test_weights"])
    problems = [trace.strip().splitlines()[-1] for _case, trace in result.failures + result.errors]
    assert not problems and result.testsRun == 4, problems


def test_reference_reader_tests_pass(reference_tests):
    """neuralmonkey/tests/test_readers.py: the tensor2tensor text reader and the string-vector reader."""
    result = reference_tests("test_readers")
    problems = [trace.strip().splitlines()[-1] for _case, trace in result.failures + result.errors]
    assert not problems and result.testsRun == 3, problems
