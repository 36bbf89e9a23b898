"""The C-ABI library loads on a CPU-only box and exports every symbol include/d2s_b200.h declares.
No compute call is made here."""
import ctypes
import os
import re

from desktop2stereo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "d2s_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d2s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_covers_header():
    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    _lib.lib()


def test_error_convention_without_gpu():
    L = _lib.lib()
    oh, ow = ctypes.c_int(), ctypes.c_int()
    assert L.d2s_sbs_out_shape(1080, 1920, _lib.FULL_SBS, 0, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (1080, 3840)
    assert L.d2s_sbs_out_shape(1080, 1920, _lib.HALF_SBS, 0, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (1080, 1920)
    assert L.d2s_sbs_out_shape(1000, 1920, _lib.FULL_TAB, 1, ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (2160, 1920)
    rc = L.d2s_sbs_out_shape(0, 0, 9, 0, ctypes.byref(oh), ctypes.byref(ow))
    assert rc != 0 and b"d2s_sbs_out_shape" in L.d2s_last_error()
    assert L.d2s_make_sbs(None, None) != 0
    assert b"sm_100a" in L.d2s_version()
