"""RGBA8 Image / ImageView and the image->tensor sampler (oracle).

Follows crates/zaru/src/image/mod.rs:45-332 (Image, ViewData, ImageView) and the
`Cnn::new` image-map closure + `sample` in crates/zaru/src/nn/mod.rs:54-73.
The sampler is NEAREST-NEIGHBOUR POINT sampling, vectorised
with float32 NumPy so each step rounds like the reference's scalar f32 code.
"""
from __future__ import annotations

import numpy as np

from .geometry import Rect, RotatedRect, Resolution, f32, rot_ccw, round_half_away

U32_MAX_F32 = f32(4294967295.0)  # u32::MAX as f32 == 4294967296.0


class Image:
    """image/mod.rs:45-180. `buf` is HxWx4 uint8, RGBA interleaved."""

    def __init__(self, buf: np.ndarray):
        assert buf.dtype == np.uint8 and buf.ndim == 3 and buf.shape[2] == 4
        self.buf = np.ascontiguousarray(buf)

    @staticmethod
    def from_rgba8(res: Resolution, data) -> "Image":
        arr = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1)
        expected = res.width * res.height * 4
        assert arr.size == expected, f"incorrect buffer size {arr.size} for {res} image (expected {expected} bytes)"
        return Image(arr.reshape(res.height, res.width, 4).copy())

    @staticmethod
    def new(width: int, height: int) -> "Image":
        return Image(np.zeros((height, width, 4), np.uint8))

    @staticmethod
    def from_rgb_array(rgb: np.ndarray) -> "Image":
        h, w, _ = rgb.shape
        buf = np.empty((h, w, 4), np.uint8)
        buf[..., :3] = rgb
        buf[..., 3] = 255
        return Image(buf)

    def width(self):
        return self.buf.shape[1]

    def height(self):
        return self.buf.shape[0]

    def resolution(self):
        return Resolution(self.width(), self.height())

    def rect(self):
        return Rect.from_top_left(0.0, 0.0, f32(self.width()), f32(self.height()))

    def view(self, rect) -> "ImageView":
        return ImageView(self, ViewData.full(self).view(rect))

    def as_view(self) -> "ImageView":
        return self.view(self.rect())

    def data(self):
        return self.buf.reshape(-1)


class ViewData:
    """image/mod.rs:187-247."""

    def __init__(self, rect: RotatedRect):
        self.rect_ = rect

    @staticmethod
    def full(image: Image):
        return ViewData(RotatedRect(image.rect(), 0.0))

    def view(self, rect):
        """image/mod.rs:201-210."""
        rect = RotatedRect.of(rect)
        radians = self.rect_.radians + rect.radians
        cx, cy = self.rect_.transform_out(rect.rect.center())
        px = cx - rect.rect.w * f32(0.5)
        py = cy - rect.rect.h * f32(0.5)
        return ViewData(RotatedRect(rect.rect.move_to(px, py), radians))

    def rect(self):
        return Rect.from_top_left(0.0, 0.0, self.width(), self.height())

    def width(self):
        return self.rect_.rect.w

    def height(self):
        return self.rect_.rect.h

    def image_coords(self, xs: np.ndarray, ys: np.ndarray, img_w: int, img_h: int):
        """Vectorised `image_coord` (image/mod.rs:224-241).

        xs, ys: uint32-valued float32 arrays of view pixel coordinates.
        Returns (ix, iy, valid) with ix/iy int64.
        """
        r = self.rect_
        px = xs.astype(np.float32) + f32(0.5)
        py = ys.astype(np.float32) + f32(0.5)
        # RotatedRect::transform_out (rect.rs:417-423)
        cx, cy = r.rect.w * f32(0.5), r.rect.h * f32(0.5)
        m00, m01, m10, m11 = rot_ccw(r.radians)
        dx = px - cx
        dy = py - cy
        rx = (f32(0.0) + m00 * dx) + m01 * dy
        ry = (f32(0.0) + m10 * dx) + m11 * dy
        tlx, tly = r.rect.top_left()
        ox = (rx + cx) + tlx
        oy = (ry + cy) + tly
        fx = round_half_away(ox - f32(0.5))
        fy = round_half_away(oy - f32(0.5))
        with np.errstate(invalid="ignore"):
            bad = (fx < 0) | (fy < 0) | (np.ceil(fx) >= U32_MAX_F32) | (np.ceil(fy) >= U32_MAX_F32)
        # NaN coordinates pass every comparison above (all false) and `NaN as u32` is 0 in Rust
        fx = np.nan_to_num(np.where(bad, f32(0.0), fx), nan=0.0)
        fy = np.nan_to_num(np.where(bad, f32(0.0), fy), nan=0.0)
        ix = round_half_away(fx).astype(np.int64)
        iy = round_half_away(fy).astype(np.int64)
        valid = ~bad & (ix < img_w) & (iy < img_h)
        return ix, iy, valid


class ImageView:
    """image/mod.rs:251-332."""

    def __init__(self, image: Image, data: ViewData):
        self.image, self.data = image, data

    def rect(self):
        return self.data.rect()

    def as_view(self):
        return self

    def view(self, rect) -> "ImageView":
        return ImageView(self.image, self.data.view(rect))

    def get(self, x: int, y: int):
        """ViewData::get (image/mod.rs:242-247): RGBA tuple, Color::NONE outside."""
        ix, iy, ok = self.data.image_coords(np.array([x], np.float32), np.array([y], np.float32),
                                            self.image.width(), self.image.height())
        if not ok[0]:
            return (0, 0, 0, 0)
        return tuple(int(v) for v in self.image.buf[iy[0], ix[0]])

    def to_image(self) -> Image:
        """image/mod.rs:314-325."""
        w = int(np.ceil(self.rect().w))
        h = int(np.ceil(self.rect().h))
        ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
        return Image(self._gather(xs, ys))

    def _gather(self, xs, ys):
        ix, iy, ok = self.data.image_coords(xs, ys, self.image.width(), self.image.height())
        ix = np.where(ok, ix, 0)
        iy = np.where(ok, iy, 0)
        px = self.image.buf[iy, ix]
        px[~ok] = 0
        return px


def sample_source_coords(view: ImageView, w: int, h: int):
    """`sample` (nn/mod.rs:54-58) for the whole [h, w] grid.

    x = round((x_out / w) * view_w) as u32 (saturating), same for y.
    """
    vw, vh = view.rect().w, view.rect().h
    u = np.arange(w, dtype=np.float32) / f32(w)
    v = np.arange(h, dtype=np.float32) / f32(h)
    sx = round_half_away(u * vw)
    sy = round_half_away(v * vh)
    # `as u32`: saturating, NaN -> 0
    sx = np.clip(np.nan_to_num(sx, nan=0.0), 0.0, U32_MAX_F32).astype(np.float32)
    sy = np.clip(np.nan_to_num(sy, nan=0.0), 0.0, U32_MAX_F32).astype(np.float32)
    ys, xs = np.meshgrid(sy, sx, indexing="ij")
    return xs, ys


def image_to_tensor(view: ImageView, w: int, h: int, lo, hi, layout="NCHW") -> np.ndarray:
    """The `image_map` closure (nn/mod.rs:63-73) + ColorMapper::map (:156-166).

    Returns float32 [1,3,h,w] (NCHW) or [1,h,w,3] (NHWC).
    """
    xs, ys = sample_source_coords(view, w, h)
    px = view._gather(xs, ys)  # [h, w, 4] u8, Color::NONE where outside
    start, end = f32(lo), f32(hi)
    adjust = (end - start) / f32(255.0)
    rgb = px[..., :3].astype(np.float32) * adjust + start  # mul then add, two roundings
    rgb = rgb.astype(np.float32)
    if layout == "NCHW":
        return np.ascontiguousarray(rgb.transpose(2, 0, 1))[None]
    return np.ascontiguousarray(rgb)[None]
