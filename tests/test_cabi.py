"""The C-ABI boundary without a GPU: libvolta_b200.so loads, exports every entry point include/volta_b200.h declares,
and the ctypes binding (volta_b200/_lib.py) covers exactly that set.  No compute calls (no device here)."""
import ctypes
import os
import re

from volta_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "volta_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"^\s*(?:int|const\s+char\s*\*)\s+(vb_\w+)\s*\(", text, re.M))


def test_header_declares_the_bound_symbols():
    names = _declared()
    assert len(names) >= 40
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m volta_b200.csrc.build`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(_declared()) if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_loads_and_reports_version_without_a_device():
    lib = _lib.load()
    assert lib.vb_version() >= 1000
    assert isinstance(_lib.last_error(), str)


def test_signatures_have_as_many_arguments_as_the_header():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))
