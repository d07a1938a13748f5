"""The C-ABI library loads and exports every symbol include/umpr_b200.h declares (no compute, CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "umpr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(umpr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    names = _declared()
    assert "umpr_gru_recurrence_fwd" in names and "umpr_coattn_fwd" in names and len(names) >= 25


def test_library_exports_every_declared_symbol():
    from umpr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.umpr_version.restype = ctypes.c_int
    assert lib.umpr_version() == 100


def test_binding_table_matches_header():
    from umpr_b200 import _lib
    declared = set(_declared()) - {"umpr_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # argument counts: every prototype in the header has as many parameters as the ctypes signature
    src = open(os.path.join(ROOT, "include", "umpr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for m in re.finditer(r"\bint\s+(umpr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name]), (name, n, len(_lib.SIGNATURES[name]))


def test_missing_library_fails_loudly(monkeypatch):
    from umpr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libumpr_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
