"""CPU: the C-ABI library builds, loads and exports every symbol include/pnb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from pointnerf_b200 import build, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pnb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pnb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    assert os.path.exists(lib.LIB_PATH)
    so = ctypes.CDLL(lib.LIB_PATH)
    decl = _declared_symbols()
    assert len(decl) >= 10
    for name in decl:
        assert hasattr(so, name), "libpnb200.so does not export %s" % name
    assert sorted(lib.SYMBOLS) == decl, "pointnerf_b200/lib.py binds a different symbol set than include/pnb200.h"


def test_struct_sizes_match_header():
    """ctypes mirrors of the POD structs have the layout the library was compiled with."""
    l = lib.load()
    for which, ty in enumerate((lib.Grid, lib.Query, lib.ShadeOpts, lib.Mlp, lib.Points)):
        assert l.pnb_struct_size(which) == ctypes.sizeof(ty), ty.__name__
    assert l.pnb_struct_size(99) == 0


def test_errors_are_reported_not_thrown():
    l = lib.load()
    assert l.pnb_version() >= 100
    rc = l.pnb_grid_build(None, None, 0, None, 0, None, None, None, None, 0, 0, 0, None, None)
    assert rc == -1
    assert b"null argument" in l.pnb_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(lib.PnbError, match="no CPU fallback"):
        lib.load()
