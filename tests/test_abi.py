"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/imagd_b200.h
declares, and the ctypes signature table matches the header (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "imagd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(imagd_\w+)\s*\(([^;{]*)\)\s*;", src)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in protos}


def test_library_exports_every_declared_symbol():
    from imagdressing_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.load()
    decl = _header_functions()
    assert len(decl) >= 17
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/imagd_b200.h but not exported"
    assert lib.imagd_version() == 100


def test_ctypes_table_matches_header():
    from imagdressing_b200 import _lib

    decl = _header_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, args in decl.items():
        assert len(args) == len(_lib.SIGNATURES[name][1]), f"{name}: header has {len(args)} args"


def test_bad_arguments_are_rejected_without_a_gpu():
    from imagdressing_b200 import _lib

    lib = _lib.load()
    rc = lib.imagd_gemm_bf16(None, 0, None, 0, None, 0, 0, 0, 0, None, None)
    assert rc == -1
    assert b"gemm" in lib.imagd_last_error()
    assert lib.imagd_groupnorm_ws_bytes(2, 4096, 320, 32) == (1024 + 2 * 64 * 32 * 2) * 4


def test_no_silent_fallback_when_library_missing(monkeypatch, tmp_path):
    from imagdressing_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ImagdError):
        _lib.load()
