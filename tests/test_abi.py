"""The C-ABI library loads and exports every symbol include/zaru_b200.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

from zaru_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "zaru_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _ffi.load_library()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in zaru_b200.h but not exported by libzaru_b200.so"
        assert n in _ffi.SIGNATURES, f"{n} has no ctypes signature in zaru_b200/_ffi.py"
    assert sorted(_ffi.SIGNATURES) == names


def test_struct_layouts():
    assert C.sizeof(_ffi.zb_view) == 24
    assert C.sizeof(_ffi.zb_detection) == 4 * (6 + 14 + 2)
    assert _ffi.lib().zb_version().decode().startswith("zaru_b200")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    st = _ffi.lib().zb_ctx_create(0, C.byref(h))
    assert st == _ffi.ZB_ERR_NO_DEVICE
    assert b"no CPU fallback" in _ffi.lib().zb_last_error()


def test_null_arguments_are_rejected_not_crashed():
    lib = _ffi.lib()
    assert lib.zb_net_load(None, None, 0, None) == _ffi.ZB_ERR_INVALID_ARGUMENT
    assert lib.zb_detector_set_threshold(None, 0.5) == _ffi.ZB_ERR_INVALID_ARGUMENT
    assert lib.zb_net_num_outputs(None) == 0
    lib.zb_net_destroy(None)
    lib.zb_frames_destroy(None)
    lib.zb_ctx_destroy(None)
