"""The C-ABI library builds, loads without a GPU driver and exports every symbol the header declares."""
import ctypes
import os
import re

from obman_train_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    protos = _lib.parse_header()
    assert len(protos) >= 10
    for name in protos:
        assert hasattr(lib, name), "missing export " + name
    with open(os.path.join(ROOT, "include", "obman_b200.h")) as f:
        declared = set(re.findall(r"\b(obman_\w+)\s*\(", f.read()))
    assert declared == set(protos), declared ^ set(protos)


def test_version_and_error_string_without_gpu():
    lib = _lib.load()
    assert lib.obman_version() >= 100
    assert isinstance(lib.obman_get_last_error(), bytes)


def test_bad_arguments_are_rejected_before_any_launch():
    lib = _lib.load()
    rc = lib.obman_nn_fwd(None, None, 0, 0, 0, None, None, None, None, 3, None)
    assert rc == -1
    assert b"obman_nn_fwd" in lib.obman_get_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "obman_train_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
