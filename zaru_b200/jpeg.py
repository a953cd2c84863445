"""Host-side mirror of `zaru_image::jpeg` (crates/zaru-image/src/jpeg.rs:107-222): `decode_jpeg(bytes)`, batched, with the
pixels landing in an uploaded ImageBatch.  Entropy decoding on the host, everything per-pixel on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .image import ImageBatch
from .rect import Resolution


def jpeg_info(data: bytes):
    """(width, height, components, h_samp, v_samp) from the header; raises ZaruError for progressive / malformed files."""
    w, h, nc, hs, vs = (C.c_int32() for _ in range(5))
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    _ffi.check(_ffi.lib().zb_jpeg_info(buf, len(data), C.byref(w), C.byref(h), C.byref(nc), C.byref(hs), C.byref(vs)))
    return w.value, h.value, nc.value, hs.value, vs.value


def jpeg_coefficients(data: bytes):
    """(coef [blocks, 64] int16, blocks_w[3], blocks_h[3], qtables [3, 64] uint16): what the host front end extracts."""
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    need = C.c_size_t()
    bw, bh = (C.c_int32 * 3)(), (C.c_int32 * 3)()
    qt = np.zeros((3, 64), np.uint16)
    _ffi.check(_ffi.lib().zb_jpeg_coefficients(buf, len(data), None, 0, C.byref(need), bw, bh, qt.ctypes.data))
    out = np.empty(need.value, np.int16)
    _ffi.check(_ffi.lib().zb_jpeg_coefficients(buf, len(data), out.ctypes.data, out.size, None, bw, bh, qt.ctypes.data))
    return out.reshape(-1, 64), list(bw), list(bh), qt


def decode_jpegs_into(batch: ImageBatch, jpegs, first: int = 0):
    """`decode_jpeg` for a list of byte strings into frames [first, first + len(jpegs)) of an uploaded batch."""
    n = len(jpegs)
    bufs = [(C.c_char * len(j)).from_buffer_copy(j) for j in jpegs]
    ptrs = (C.c_void_p * n)(*[C.addressof(b) for b in bufs])
    sizes = (C.c_size_t * n)(*[len(j) for j in jpegs])
    _ffi.check(_ffi.lib().zb_frames_decode_jpeg(batch._h, first, ptrs, sizes, n))


def decode_jpeg(data: bytes) -> np.ndarray:
    """One JPEG -> RGBA8 [h, w, 4] (decoded on the device, read back)."""
    w, h, *_ = jpeg_info(data)
    batch = ImageBatch.from_rgba8(Resolution(w, h), np.zeros((1, h, w, 4), np.uint8))
    decode_jpegs_into(batch, [data])
    return batch.frame(0).as_view().to_image()._pixels
