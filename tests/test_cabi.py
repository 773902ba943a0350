"""The C-ABI shared library: builds, loads without a GPU, exports every declared symbol."""
import ctypes
import os
import re

from neuralmonkey_b200 import lib


def test_library_loads_and_reports_version():
    handle = lib.load()
    assert handle.nm_version() == 1
    assert handle.nm_last_error() is not None


def test_every_declared_symbol_is_exported():
    handle = ctypes.CDLL(lib.LIB_PATH)
    declared = lib.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(handle, name), "include/nmb200.h declares {} but the .so lacks it".format(name)


def test_header_cites_reference_for_every_entry_point():
    text = open(lib.HEADER_PATH, encoding="utf-8").read()
    # every K-section names the reference file it replaces
    for needle in ("model/sequence.py", "encoders/recurrent.py", "attention/feed_forward.py",
                   "decoders/autoregressive.py", "decoders/beam_search_decoder.py",
                   "trainers/generic_trainer.py", "attention/scaled_dot_product.py",
                   "encoders/imagenet_encoder.py", "tf_utils.py"):
        assert needle in text, needle
    code = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)  # declarations only, comments stripped
    assert not re.search(r"torch|at::|Tensor|std::", code), "no torch / C++ types in the C ABI"


def test_invalid_arguments_are_reported_not_crashed():
    handle = lib.load()
    rc = handle.nm_gemm(0, 0, 4, 4, 4, None, 4, None, 4, None, 4, None, 0, 0.0, 0, None)
    assert rc < 0
    assert b"null" in handle.nm_last_error()
    rc = handle.nm_gemm_uses_tc(0, 0, 128, 128, 30, 30, 128, 128)
    assert rc == 0  # K=30 rows are not 16-byte multiples
    assert handle.nm_gemm_uses_tc(0, 0, 128, 128, 32, 32, 128, 128) == 1


def test_product_code_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "neuralmonkey_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
