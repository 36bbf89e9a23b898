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


def test_jpeg_sizing_calls_without_gpu():
    """d2s_jpeg_workspace_bytes / d2s_jpeg_max_bytes are host arithmetic: geometry, monotonicity and the error convention"""
    L = _lib.lib()
    # 1080p Full-SBS, 4 MCUs per restart interval: 240 x 68 MCUs -> 4080 intervals
    ws = L.d2s_jpeg_workspace_bytes(1080, 3840, 4)
    coef = 240 * 68 * 6 * 64 * 2
    slots = 4080 * (4 * 6 * 416 + 16)
    assert ws >= coef + slots + 2 * 4080 * 4 and ws < coef + slots + (1 << 20)
    assert L.d2s_jpeg_max_bytes(1080, 3840, 4) >= 1080 * 3840 * 3          # can never overflow: above the raw frame
    assert L.d2s_jpeg_workspace_bytes(2160, 7680, 4) > ws
    for bad in ((1081, 3840, 4), (1080, 3841, 4), (1080, 3840, 0), (1080, 3840, 70000), (0, 0, 4)):
        assert L.d2s_jpeg_workspace_bytes(*bad) == 0 and L.d2s_jpeg_max_bytes(*bad) == 0
        assert b"d2s_jpeg" in L.d2s_last_error()
    assert L.d2s_jpeg_encode(None, 0, 16, 16, 90, 1, None, 0, None, None, 0, None) != 0
    assert L.d2s_pipe_set_fps_text(None, b"FPS: 60.0") != 0
