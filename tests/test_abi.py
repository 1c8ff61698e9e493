"""The C-ABI library loads and exports exactly what include/nesie_b200.h declares (CPU-only:
no compute call is made; argument validation happens before any CUDA call)."""
import ctypes
import os
import re

import pytest

from nesie_b200 import _lib
from nesie_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "nesie_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nesie_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build()
    assert os.path.exists(path)
    lib = _lib.lib()
    assert lib.nesie_abi_version() == 1


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(build())
    names = header_functions()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_table_matches_header():
    declared = set(header_functions()) - {"nesie_last_error"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_invalid_arguments_return_status_not_exit():
    lib = _lib.lib()
    rc = lib.nesie_fps(1, 10, 2, None, None, None, None)
    assert rc == 10001
    assert b"null pointer" in lib.nesie_last_error()
    rc = lib.nesie_ball_query(1, -1, 1, 0.0, 1.0, 4, None, None, None, None)
    assert rc == 10001
    assert lib.nesie_aligned_3d_nms_batched(1, 5000, None, None, None, None, 0.25, None, None,
                                            None) == 10001
    with pytest.raises(RuntimeError):
        _lib.check(rc, "nesie_ball_query")


def test_fps_needs_temp_threshold():
    lib = _lib.lib()
    assert lib.nesie_fps_needs_temp(8, 40000, 2048) == 0
    assert lib.nesie_fps_needs_temp(16, 100000, 4096) == 0
    assert lib.nesie_fps_needs_temp(1, 200000, 16) == 1
