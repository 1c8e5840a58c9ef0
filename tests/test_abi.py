"""The C-ABI shared library loads without a GPU and exports exactly what include/mlic_b200.h declares."""
import ctypes as C
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "mlic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mlic_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib_built):
    from mlic_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_library_exports_every_symbol(lib_built):
    lib = C.CDLL(lib_built)
    for name in _declared():
        assert hasattr(lib, name), name


def test_error_reporting_without_compute(lib_built):
    from mlic_b200 import _lib
    L = _lib.lib()
    assert L.mlic_version().startswith(b"mlic_b200")
    h = C.c_void_p()
    assert L.mlic_engine_create(192, 320, 7, 0, C.byref(h)) != 0              # M % slice_num != 0 (mlicpp.py:21)
    assert b"divisible" in L.mlic_last_error()
    assert L.mlic_engine_create(192, 320, 10, 9, C.byref(h)) != 0
    assert L.mlic_engine_create(192, 320, 10, 0, C.byref(h)) == 0
    n = C.c_size_t()
    assert L.mlic_workspace_bytes(h, 0, 1, 1, 64, 64, C.byref(n)) != 0        # not finalized
    assert b"finalized" in L.mlic_last_error()
    assert L.mlic_engine_set_option(h, b"no_such_option", 1) != 0
    L.mlic_engine_destroy(h)
