"""zaru_b200 — host-side mirror of Zaru's perception API over libzaru_b200.so (CUDA, sm_100a).

Module layout follows the reference crate (`zaru::image`, `zaru::rect`, `zaru::nn`,
`zaru::detection`, `zaru::landmark`, `zaru::face`, `zaru::hand`).  Everything that computes runs in
the CUDA library; importing this package never falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _ffi
from ._ffi import ZaruError, load_library  # noqa: F401

_ctx = None
_tls = threading.local()
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def context(device: int | None = None):
    """The process-wide `zb_ctx` (one per GPU; LOCAL_RANK picks the device under torchrun)."""
    global _ctx
    own = getattr(_tls, "ctx", None)
    if own is not None:
        return own
    if _ctx is None:
        if device is None:
            device = _default_device()
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_ctx_create(device, C.byref(h)))
        _ctx = h
    return _ctx


def _default_device() -> int:
    return int(os.environ.get("ZARU_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def thread_context(device: int | None = None):
    """Give the CALLING thread its own `zb_ctx` (stream + event pool) on the GPU; everything this thread creates
    afterwards (networks, detectors, pipelines, frame batches) lives on it.  This is the reference's threading model
    (SURVEY 8b): `Detector::detect` takes `&mut self`, so parallel callers hold one detector per thread (rayon
    `map_init`, eval_face_recognition.rs:67-70) - here two threads on one GPU let one pipeline's PCIe ingest overlap
    the other's compute.  Idempotent per thread."""
    own = getattr(_tls, "ctx", None)
    if own is None:
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_ctx_create(_default_device() if device is None else device, C.byref(h)))
        _tls.ctx = own = h
    return own


def context_key() -> int:
    """Identity of the calling thread's context (network caches are per context)."""
    return int(context().value or 0)


def model_dir() -> str:
    """Where the MediaPipe `.onnx` blobs live (the reference embeds them with include_blob!,
    crates/zaru/src/face/detection.rs:38).  build() stages them under assets/_ref/onnx."""
    env = os.environ.get("ZARU_B200_MODEL_DIR")
    if env:
        return env
    staged = os.path.join(_ROOT, "assets", "_ref", "onnx")
    if os.path.isdir(staged):
        return staged
    return "/root/reference/3rdparty/onnx"


def model_path(name: str) -> str:
    return os.path.join(model_dir(), name)


def launch_count() -> int:
    return int(_ffi.lib().zb_launch_count(context()))


def last_device_ms() -> float:
    return float(_ffi.lib().zb_last_device_ms(context()))


def sync():
    _ffi.check(_ffi.lib().zb_sync(context()))


def timer_start():
    _ffi.check(_ffi.lib().zb_timer_start(context()))


def timer_stop_ms() -> float:
    ms = C.c_float()
    _ffi.check(_ffi.lib().zb_timer_stop(context(), C.byref(ms)))
    return float(ms.value)


def last_h2d_bytes() -> int:
    """Bytes the last `zb_frames_decode_jpeg` call on this thread's context sent to the device."""
    return int(_ffi.lib().zb_last_h2d_bytes(context()))


def profile_begin(per_layer: bool = False):
    """Bracket every kernel launch on this thread's context with CUDA events; per_layer: one row per network layer."""
    _ffi.check(_ffi.lib().zb_profile_set_detail(context(), 1 if per_layer else 0))
    _ffi.check(_ffi.lib().zb_profile_begin(context()))


def profile_end() -> dict:
    """Stop profiling; {row: {launches, ms, bytes, flops, kernels: {function: {...}}}} (CUDA events per launch)."""
    import json
    need = C.c_size_t(0)
    buf = C.create_string_buffer(1 << 18)
    _ffi.check(_ffi.lib().zb_profile_end(context(), buf, len(buf), C.byref(need)))
    if need.value > len(buf):
        raise RuntimeError("profile JSON truncated")
    return json.loads(buf.value.decode())
