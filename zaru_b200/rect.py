"""Host-side mirror of `zaru_image::{Rect, RotatedRect, Resolution, AspectRatio}`.

Same names and argument meaning as crates/zaru-image/src/rect.rs and resolution.rs; arithmetic is
float32 (struct-packed through `_f`) so that view rectangles handed to the CUDA library are the
ones the reference would compute.  Only the operations on the perception path are mirrored.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

_f = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cosf.restype = _libm.sinf.restype = ctypes.c_float
_libm.cosf.argtypes = _libm.sinf.argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]


def signed_angle_to(ax, ay, bx, by):
    """`Vec2::signed_angle_to` (zaru-linalg vector.rs:568-573): -perp_dot(a, b).atan2(dot(a, b)), f32."""
    ax, ay, bx, by = _f(ax), _f(ay), _f(bx), _f(by)
    perp = ax * by - ay * bx
    dot = (_f(0.0) + ax * bx) + ay * by
    return -_f(_libm.atan2f(float(perp), float(dot)))


def _cos_sin(radians):
    return _f(_libm.cosf(float(radians))), _f(_libm.sinf(float(radians)))


def _rot(c, s, x, y):
    # Mat2::rotation_counterclockwise * v, folded from 0 (zaru-linalg matrix.rs:571-579, matrix/ops.rs:74-76)
    return (_f(0.0) + c * x) + (-s) * y, (_f(0.0) + s * x) + c * y


class AspectRatio:
    def __init__(self, width: int, height: int):
        if width == 0 or height == 0:
            raise ValueError("AspectRatio needs non-zero width and height")
        g = math.gcd(int(width), int(height))
        self.width, self.height = int(width) // g, int(height) // g

    def as_f32(self):
        return _f(self.width) / _f(self.height)

    def __eq__(self, other):
        return (self.width, self.height) == (other.width, other.height)

    def __repr__(self):
        return f"{self.width}:{self.height}"


AspectRatio.SQUARE = AspectRatio(1, 1)


class Resolution:
    def __init__(self, width: int, height: int):
        self._w, self._h = int(width), int(height)

    def width(self):
        return self._w

    def height(self):
        return self._h

    def num_pixels(self):
        return self._w * self._h

    def aspect_ratio(self):
        return None if self._w == 0 or self._h == 0 else AspectRatio(self._w, self._h)

    def __eq__(self, other):
        return (self._w, self._h) == (other._w, other._h)

    def __repr__(self):
        return f"{self._w}x{self._h}"


Resolution.RES_1080P = Resolution(1920, 1080)
Resolution.RES_720P = Resolution(1280, 720)


class Rect:
    """Axis-aligned rectangle stored as centre + size (rect.rs:14-18)."""

    __slots__ = ("_cx", "_cy", "_w", "_h")

    def __init__(self, cx, cy, w, h):
        self._cx, self._cy, self._w, self._h = _f(cx), _f(cy), _f(w), _f(h)

    @classmethod
    def from_center(cls, x_center, y_center, width, height):
        return cls(x_center, y_center, width, height)

    @classmethod
    def from_top_left(cls, x, y, width, height):
        x, y, width, height = _f(x), _f(y), _f(width), _f(height)
        return cls(x + width * _f(0.5), y + height * _f(0.5), width, height)

    def center(self):
        return (self._cx, self._cy)

    def size(self):
        return (self._w, self._h)

    def width(self):
        return self._w

    def height(self):
        return self._h

    def top_left(self):
        return (self._cx - self._w * _f(0.5), self._cy - self._h * _f(0.5))

    def x(self):
        return self.top_left()[0]

    def y(self):
        return self.top_left()[1]

    def area(self):
        return self._w * self._h

    def move_to(self, x, y):
        return Rect.from_top_left(x, y, self._w, self._h)

    def move_by(self, offset):
        return Rect(self._cx + _f(offset[0]), self._cy + _f(offset[1]), self._w, self._h)

    @classmethod
    def bounding(cls, points):
        """`Rect::bounding` (rect.rs:49-68); None for no points."""
        pts = [(_f(p[0]), _f(p[1])) for p in points]
        if not pts:
            return None
        minx, miny = pts[0]
        maxx, maxy = pts[0]
        for x, y in pts[1:]:
            minx, miny, maxx, maxy = min(minx, x), min(miny, y), max(maxx, x), max(maxy, y)
        return cls.from_top_left(minx, miny, maxx - minx, maxy - miny)

    def scale(self, s):
        return Rect(self._cx, self._cy, self._w * _f(s), self._h * _f(s))

    def grow_rel(self, amount):
        a = _f(amount)
        l, r, t, b = self._w * a, self._w * a, self._h * a, self._h * a
        return Rect(self._cx, self._cy, self._w + l + r, self._h + t + b)

    def grow_to_fit_aspect(self, target: AspectRatio):
        w, h = self._w, self._h
        target_width = h * target.as_f32()
        if target_width >= w:
            w = w + (target_width - w)
        else:
            h = h + (w / target.as_f32() - h)
        return Rect(self._cx, self._cy, w, h)

    def __eq__(self, other):
        return (self._cx, self._cy, self._w, self._h) == (other._cx, other._cy, other._w, other._h)

    def __repr__(self):
        return f"Rect @ ({self._cx},{self._cy})/{self._w}x{self._h}"


class RotatedRect:
    """A Rect rotated clockwise around its centre (rect.rs:269-273)."""

    __slots__ = ("_rect", "_radians")

    def __init__(self, rect: Rect, radians=0.0):
        self._rect, self._radians = rect, _f(radians)

    @classmethod
    def new(cls, rect, radians):
        return cls(rect, radians)

    @classmethod
    def of(cls, r):
        return r if isinstance(r, RotatedRect) else cls(r, 0.0)

    @classmethod
    def bounding(cls, radians, points):
        """`RotatedRect::bounding(radians, points)` (rect.rs:287-325); None for no points."""
        pts = [(_f(p[0]), _f(p[1])) for p in points]
        if not pts:
            return None
        radians = _f(radians)
        c, s = _cos_sin(-radians)                   # rotation_clockwise(r) = rotation_counterclockwise(-r)
        fmax = _f(np.finfo(np.float32).max)
        minx = miny = fmax
        maxx = maxy = -fmax
        for x, y in pts:
            px, py = _rot(c, s, x, y)
            minx, miny, maxx, maxy = min(minx, px), min(miny, py), max(maxx, px), max(maxy, py)
        ccx, ccy = (minx + maxx) * _f(0.5), (miny + maxy) * _f(0.5)
        c2, s2 = _cos_sin(radians)
        cx, cy = _rot(c2, s2, ccx, ccy)             # center.rotate_counterclockwise(radians)
        return cls(Rect.from_center(cx, cy, maxx - minx, maxy - miny), radians)

    def rect(self):
        return self._rect

    def rotation_radians(self):
        return self._radians

    def center(self):
        return self._rect.center()

    def map(self, f):
        return RotatedRect(f(self._rect), self._radians)

    def grow_rel(self, amount):
        return self.map(lambda r: r.grow_rel(amount))

    def grow_to_fit_aspect(self, target):
        return self.map(lambda r: r.grow_to_fit_aspect(target))

    def transform_out(self, pt):
        hx, hy = self._rect._w * _f(0.5), self._rect._h * _f(0.5)
        c, s = _cos_sin(self._radians)
        rx, ry = _rot(c, s, _f(pt[0]) - hx, _f(pt[1]) - hy)
        tlx, tly = self._rect.top_left()
        return (rx + hx + tlx, ry + hy + tly)

    def transform_in(self, pt):
        hx, hy = self._rect._w * _f(0.5), self._rect._h * _f(0.5)
        tlx, tly = self._rect.top_left()
        c, s = _cos_sin(-self._radians)
        rx, ry = _rot(c, s, _f(pt[0]) - tlx - hx, _f(pt[1]) - tly - hy)
        return (rx + hx, ry + hy)

    def __eq__(self, other):
        return self._rect == other._rect and self._radians == other._radians

    def __repr__(self):
        return f"RotatedRect({self._rect!r}, {self._radians})"
