"""ctypes binding of libzaru_b200.so — one Python declaration per symbol of include/zaru_b200.h.

The library is the product; this module only loads it.  There is NO CPU fallback: if the shared
library is missing or no CUDA device is present, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzaru_b200.so")

ZB_OK = 0
ZB_ERR_INVALID_ARGUMENT = -1
ZB_ERR_CUDA = -2
ZB_ERR_BAD_MODEL = -3
ZB_ERR_UNSUPPORTED_OP = -4
ZB_ERR_BAD_SHAPE = -5
ZB_ERR_CAPACITY = -6
ZB_ERR_NO_DEVICE = -7

ZB_DET_FACE_SHORT_RANGE, ZB_DET_PALM, ZB_DET_FACE_FULL_RANGE = 0, 1, 2
ZB_EST_FACE_MESH_V1, ZB_EST_EYE, ZB_EST_HAND, ZB_EST_FACE_MESH_V2 = 0, 1, 2, 3
ZB_NMS_REMOVE, ZB_NMS_AVERAGE = 0, 1
ZB_NCHW, ZB_NHWC = 0, 1
ZB_MAX_KEYPOINTS = 7


class zb_view(C.Structure):
    _fields_ = [("frame", C.c_int32), ("cx", C.c_float), ("cy", C.c_float), ("w", C.c_float), ("h", C.c_float),
                ("radians", C.c_float)]


class zb_detection(C.Structure):
    _fields_ = [("confidence", C.c_float), ("angle", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("w", C.c_float), ("h", C.c_float), ("keypoints", C.c_float * (2 * ZB_MAX_KEYPOINTS)),
                ("num_keypoints", C.c_int32), ("anchor", C.c_int32)]


P = C.c_void_p
PP = C.POINTER(C.c_void_p)
i32, i64, f32, sz = C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes).  Keep in sync with include/zaru_b200.h (tests/test_abi.py checks it).
SIGNATURES = {
    "zb_last_error": (C.c_char_p, []),
    "zb_version": (C.c_char_p, []),
    "zb_ctx_create": (i32, [i32, PP]),
    "zb_ctx_destroy": (None, [P]),
    "zb_sync": (i32, [P]),
    "zb_launch_count": (i64, [P]),
    "zb_net_load": (i32, [P, P, sz, PP]),
    "zb_net_destroy": (None, [P]),
    "zb_net_num_inputs": (i32, [P]),
    "zb_net_num_outputs": (i32, [P]),
    "zb_net_input_info": (i32, [P, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i64)]),
    "zb_net_output_info": (i32, [P, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i64)]),
    "zb_net_estimate": (i32, [P, P, i32, C.POINTER(C.c_void_p)]),
    "zb_net_set_chunk": (i32, [P, i32]),
    "zb_frames_upload": (i32, [P, P, i32, i32, i64, i32, PP]),
    "zb_frames_alias": (i32, [P, P, i32, i32, i64, i32, PP]),
    "zb_frames_update": (i32, [P, P, i32, i32]),
    "zb_frames_destroy": (None, [P]),
    "zb_preprocess": (i32, [P, P, P, i32, i32, i32, f32, f32, i32, P]),
    "zb_detector_create": (i32, [P, P, i32, f32, f32, PP]),
    "zb_detector_destroy": (None, [P]),
    "zb_detector_set_threshold": (i32, [P, f32]),
    "zb_detector_set_nms": (i32, [P, f32, i32]),
    "zb_detector_input_resolution": (i32, [P, C.POINTER(i32), C.POINTER(i32)]),
    "zb_detector_detect": (i32, [P, P, P, i32, P, P, i32, P, P]),
    "zb_detector_extract": (i32, [P, P, P, P, i32, P, P, i32]),
    "zb_estimator_create": (i32, [P, P, i32, f32, f32, PP]),
    "zb_estimator_destroy": (None, [P]),
    "zb_estimator_num_landmarks": (i32, [P]),
    "zb_estimator_input_resolution": (i32, [P, C.POINTER(i32), C.POINTER(i32)]),
    "zb_estimator_estimate": (i32, [P, P, P, P, i32, P, P]),
    "zb_face_pipeline_create": (i32, [P, P, P, PP]),
    "zb_face_pipeline_destroy": (None, [P]),
    "zb_tracker_create": (i32, [P, P, i32, f32, f32, i32, PP]),
    "zb_tracker_destroy": (None, [P]),
    "zb_tracker_set_loss_threshold": (i32, [P, f32]),
    "zb_tracker_set_roi_padding": (i32, [P, f32]),
    "zb_tracker_set_roi": (i32, [P, P, P, i32]),
    "zb_tracker_roi": (i32, [P, P, P]),
    "zb_tracker_track": (i32, [P, P, i32, P, P, P, P, P]),
    "zb_estimator_set_filter": (i32, [P, i32, f32, f32, f32, f32]),
    "zb_tracker_set_filter": (i32, [P, i32, f32, f32, f32, f32]),
    "zb_filter_apply": (i32, [P, i32, f32, f32, f32, f32, P, P, i64]),
    "zb_frames_decode_jpeg": (i32, [P, i32, P, P, i32]),
    "zb_jpeg_info": (i32, [P, sz, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "zb_jpeg_coefficients": (i32, [P, sz, P, sz, C.POINTER(sz), C.POINTER(i32), C.POINTER(i32), P]),
    "zb_last_h2d_bytes": (i64, [P]),
    "zb_blend": (i32, [P, P, P, P, P, i32]),
    "zb_view_to_image": (i32, [P, P, P, i32, i32, i32, P]),
    "zb_frames_clear": (i32, [P, i32, i32, P]),
    "zb_hand_pipeline_create": (i32, [P, P, P, PP]),
    "zb_hand_pipeline_destroy": (None, [P]),
    "zb_hand_pipeline_set_threshold": (i32, [P, f32, f32, i32]),
    "zb_hand_pipeline_run": (i32, [P, P, i32, P, P, i32, P, P, P]),
    "zb_face_pipeline_set_threshold": (i32, [P, f32, f32, i32]),
    "zb_face_pipeline_set_dense": (i32, [P, i32]),
    "zb_hand_pipeline_set_dense": (i32, [P, i32]),
    "zb_face_pipeline_num_landmarks": (i32, [P]),
    "zb_face_pipeline_run": (i32, [P, P, i32, P, P, i32, P, P, P]),
    "zb_detector_timers": (i32, [P, C.POINTER(f32)]),
    "zb_estimator_timers": (i32, [P, C.POINTER(f32)]),
    "zb_face_iris_pipeline_create": (i32, [P, P, P, PP]),
    "zb_face_iris_pipeline_destroy": (None, [P]),
    "zb_face_iris_pipeline_set_eye_margin": (i32, [P, f32]),
    "zb_face_iris_pipeline_num_landmarks": (i32, [P]),
    "zb_face_iris_pipeline_run": (i32, [P, P, P, i32, P, P, P, P, P]),
    "zb_net_plan_json": (i32, [P, C.c_char_p, sz, C.POINTER(sz)]),
    "zb_net_weights": (i32, [P, C.POINTER(C.POINTER(C.c_float)), C.POINTER(sz)]),
    "zb_plan_from_onnx": (i32, [P, sz, i32, C.c_char_p, sz, C.POINTER(sz), P, sz, C.POINTER(sz)]),
    "zb_debug_tc_gemm": (i32, [P, P, P, P, i32, i32, i32]),
    "zb_debug_mma_rate": (i32, [P, i32, i32, i32, i32, i32, i32, i32, P]),
    "zb_last_device_ms": (f32, [P]),
    "zb_host_alloc": (i32, [C.c_size_t, PP]),
    "zb_host_free": (None, [P]),
    "zb_timer_start": (i32, [P]),
    "zb_timer_stop": (i32, [P, C.POINTER(f32)]),
    "zb_profile_begin": (i32, [P]),
    "zb_profile_end": (i32, [P, C.c_char_p, sz, C.POINTER(sz)]),
    "zb_profile_set_detail": (i32, [P, i32]),
}

_lib = None


class ZaruError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"zaru_b200 error {status}: {message}")
        self.status = status
        self.message = message


def load_library(path: str | None = None):
    """dlopen the in-tree CUDA library and declare every entry point. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("ZARU_B200_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "zaru_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def lib():
    return load_library()


def check(status: int):
    if status != ZB_OK:
        raise ZaruError(status, lib().zb_last_error().decode(errors="replace"))
