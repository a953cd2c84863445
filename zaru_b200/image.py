"""Host-side mirror of `zaru::image::{Image, ImageView, AsImageView}` (crates/zaru/src/image/mod.rs).

An `Image` owns RGBA8 pixels on the host and, lazily, a copy in HBM (a one-frame `zb_frames`
batch); `ImageBatch` holds n same-sized frames resident on the device.  `ImageView` only carries
the composed view rectangle (`ViewData`, image/mod.rs:187-210); pixels are never touched on the
host — sampling happens in the CUDA `sample_kernel`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .rect import Rect, Resolution, RotatedRect, _f


def _ctx():
    from . import context
    return context()


class ImageBatch:
    """n RGBA8 frames of identical size in HBM (`zb_frames`)."""

    def __init__(self, handle, width, height, n, keepalive=None):
        self._h, self._w, self._hgt, self._n, self._keep = handle, width, height, n, keepalive

    @classmethod
    def from_rgba8(cls, res: Resolution, frames: np.ndarray) -> "ImageBatch":
        """frames: uint8 [n, height, width, 4] host array (copied to the device)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        if frames.ndim != 4 or frames.shape[1:] != (res.height(), res.width(), 4):
            raise ValueError(f"incorrect buffer shape {frames.shape} for {res} frames")
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_frames_upload(_ctx(), frames.ctypes.data, res.width(), res.height(),
                                               res.width() * 4, frames.shape[0], C.byref(h)))
        return cls(h, res.width(), res.height(), frames.shape[0])

    @classmethod
    def alias_device(cls, res: Resolution, device_ptr: int, n: int, keepalive=None) -> "ImageBatch":
        """Wrap frames that already live in HBM (e.g. a torch uint8 CUDA tensor's data_ptr())."""
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_frames_alias(_ctx(), C.c_void_p(device_ptr), res.width(), res.height(),
                                              res.width() * 4, n, C.byref(h)))
        return cls(h, res.width(), res.height(), n, keepalive)

    @classmethod
    def alias_pinned_host(cls, res: Resolution, host_ptr: int, n: int, keepalive=None) -> "ImageBatch":
        """Frames stay in PINNED host memory; the CUDA sampler reads the texels it needs across PCIe."""
        return cls.alias_device(res, host_ptr, n, keepalive)

    def update(self, frames: np.ndarray, first: int = 0):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        _ffi.check(_ffi.lib().zb_frames_update(self._h, frames.ctypes.data, first, frames.shape[0]))

    def clear(self, color=(0, 0, 0, 0), first: int = 0, count: int = None):
        """`Image::clear(color)` (image/mod.rs:171-173) on frames [first, first + count) of an uploaded batch."""
        c = (C.c_uint8 * 4)(*[int(v) for v in color])
        _ffi.check(_ffi.lib().zb_frames_clear(self._h, first, self._n - first if count is None else count, c))

    def __len__(self):
        return self._n

    def resolution(self):
        return Resolution(self._w, self._hgt)

    def frame(self, index: int) -> "Image":
        return Image(None, self, index)

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_frames_destroy(self._h)
                self._h = None
        except Exception:
            pass


class Image:
    """An 8-bit sRGB image with alpha channel (image/mod.rs:45-51)."""

    def __init__(self, pixels, batch: ImageBatch | None = None, index: int = 0):
        self._pixels = pixels
        self._batch, self._index = batch, index
        if batch is not None:
            self._w, self._h = batch._w, batch._hgt
        else:
            self._h, self._w = pixels.shape[:2]

    @classmethod
    def from_rgba8(cls, res: Resolution, buf) -> "Image":
        arr = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.reshape(-1)
        expected = res.width() * res.height() * 4
        if arr.size != expected:
            raise ValueError(f"incorrect buffer size {arr.size} for {res} image (expected {expected} bytes)")
        return cls(np.array(arr, dtype=np.uint8).reshape(res.height(), res.width(), 4))

    @classmethod
    def new(cls, width: int, height: int) -> "Image":
        return cls(np.zeros((height, width, 4), np.uint8))

    def width(self):
        return self._w

    def height(self):
        return self._h

    def resolution(self):
        return Resolution(self._w, self._h)

    def rect(self) -> Rect:
        return Rect.from_top_left(0.0, 0.0, _f(self._w), _f(self._h))

    def data(self):
        return self._pixels.reshape(-1)

    def device(self):
        """(batch, frame index) of this image's HBM copy (uploaded on first use)."""
        if self._batch is None:
            self._batch = ImageBatch.from_rgba8(self.resolution(), self._pixels[None])
            self._index = 0
        return self._batch, self._index

    def view(self, rect) -> "ImageView":
        return ImageView(self, _full_view_data(self)).view(rect)

    def as_view(self) -> "ImageView":
        return self.view(self.rect())


def _full_view_data(image: Image) -> RotatedRect:
    return RotatedRect(image.rect(), 0.0)


class ImageView:
    """An immutable view of a rectangular (possibly rotated, possibly oversized) section of an Image."""

    def __init__(self, image: Image, data: RotatedRect):
        self._image, self._data = image, data

    def rect(self) -> Rect:
        r = self._data.rect()
        return Rect.from_top_left(0.0, 0.0, r.width(), r.height())

    def view(self, rect) -> "ImageView":
        # ViewData::view (image/mod.rs:201-210)
        rect = RotatedRect.of(rect)
        radians = self._data.rotation_radians() + rect.rotation_radians()
        cx, cy = self._data.transform_out(rect.rect().center())
        w, h = rect.rect().size()
        pos = (cx - w * _f(0.5), cy - h * _f(0.5))
        return ImageView(self._image, RotatedRect(rect.rect().move_to(pos[0], pos[1]), radians))

    def as_view(self) -> "ImageView":
        return self

    def to_image(self) -> "Image":
        """`ImageView::to_image` (image/mod.rs:314-325): the view's pixels (nearest texel through the rotation,
        Color::NONE outside the image) as a new Image of size ceil(width) x ceil(height); sampled on the device."""
        import math
        r = self.rect()
        w, h = int(math.ceil(float(r.width()))), int(math.ceil(float(r.height())))
        batch, idx = self._image.device()
        out = np.empty((h, w, 4), np.uint8)
        view = (_ffi.zb_view * 1)(self.to_zb_view(idx))
        from . import context
        _ffi.check(_ffi.lib().zb_view_to_image(context(), batch._h, view, 1, w, h, out.ctypes.data))
        return Image(out)

    def image(self) -> Image:
        return self._image

    def view_rect(self) -> RotatedRect:
        """The composed rectangle in root-image coordinates (`ViewData::rect`)."""
        return self._data

    def to_zb_view(self, frame_index: int) -> _ffi.zb_view:
        r = self._data.rect()
        cx, cy = r.center()
        return _ffi.zb_view(frame_index, float(cx), float(cy), float(r.width()), float(r.height()),
                            float(self._data.rotation_radians()))

    def __repr__(self):
        return f"ImageView @ {self._data!r}"


def blend(dest, src):
    """`zaru_image::blend(&mut dest, &src)` (crates/zaru-image/src/blend.rs:13-32): draws the source view over the destination
    view with linear filtering; scaled up or down as the view sizes demand; source UVs outside the image write Color::NONE.
    `dest` / `src`: Image or ImageView whose images live in uploaded ImageBatches (the destination is modified on the
    device; read it back with `.to_image()`).  The reference performs the operation when the returned `BlendOp` is dropped;
    here it happens at once."""
    from . import context
    dv, sv = dest.as_view(), src.as_view()
    dbatch, didx = dv._image.device()
    sbatch, sidx = sv._image.device()
    d = (_ffi.zb_view * 1)(dv.to_zb_view(didx))
    s = (_ffi.zb_view * 1)(sv.to_zb_view(sidx))
    _ffi.check(_ffi.lib().zb_blend(context(), dbatch._h, d, sbatch._h, s, 1))
