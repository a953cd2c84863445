"""Rect / RotatedRect / Resolution / AspectRatio restated in float32 (oracle).

Follows crates/zaru-image/src/rect.rs and resolution.rs and the used subset of
crates/zaru-linalg (vector.rs:514-597, matrix.rs:563-579, matrix/ops.rs:74-76).
All arithmetic is done on np.float32 scalars so every operation rounds exactly
like Rust's f32.  Transcendentals go through glibc's libm (cosf/sinf/atan2f/
expf), which is what Rust's f32::cos/sin/atan2/exp call on linux-gnu.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math

import numpy as np

f32 = np.float32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "expf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]


def cosf(x):
    return f32(_libm.cosf(float(x)))


def sinf(x):
    return f32(_libm.sinf(float(x)))


def expf(x):
    return f32(_libm.expf(float(x)))


def atan2f(y, x):
    return f32(_libm.atan2f(float(y), float(x)))


def round_half_away(x):
    """Rust f32::round (half away from zero) for scalars or arrays of float32."""
    x = np.asarray(x, dtype=np.float32)
    t = np.trunc(x)
    frac = x - t  # exact
    adj = np.where(np.abs(frac) >= f32(0.5), np.copysign(f32(1.0), x), f32(0.0)).astype(np.float32)
    return (t + adj).astype(np.float32)


def sigmoid(v):
    """crates/zaru/src/num.rs:6-8: 1.0 / (1.0 + (-v).exp())."""
    return f32(1.0) / (f32(1.0) + expf(-f32(v)))


def rot_ccw(radians):
    """Mat2f::rotation_counterclockwise (matrix.rs:571-579) -> rows ((c,-s),(s,c))."""
    r = f32(radians)
    c, s = cosf(r), sinf(r)
    return (c, -s, s, c)


def mat_mul_vec(m, x, y):
    """Matrix*Vector fold: (0 + m_r0*x) + m_r1*y (matrix/ops.rs:74-76)."""
    m00, m01, m10, m11 = m
    rx = (f32(0.0) + m00 * x) + m01 * y
    ry = (f32(0.0) + m10 * x) + m11 * y
    return rx, ry


def rotate_ccw(x, y, radians):
    return mat_mul_vec(rot_ccw(radians), f32(x), f32(y))


def rotate_cw(x, y, radians):
    # rotation_clockwise(r) = rotation_counterclockwise(-r) (matrix.rs:563-569)
    return mat_mul_vec(rot_ccw(-f32(radians)), f32(x), f32(y))


def signed_angle_to(ax, ay, bx, by):
    """Vec2::signed_angle_to (vector.rs:568-573): -atan2(perp_dot, dot).

    perp_dot = a1*b2 - a2*b1 (cross z, vector.rs:645-657);
    dot = fold(0, acc + a*b) (vector.rs:350-358).
    """
    ax, ay, bx, by = f32(ax), f32(ay), f32(bx), f32(by)
    perp = ax * by - ay * bx
    dot = (f32(0.0) + ax * bx) + ay * by
    return -atan2f(perp, dot)


class AspectRatio:
    """resolution.rs:128-162 (gcd-reduced width:height)."""

    def __init__(self, width: int, height: int):
        if width == 0 or height == 0:
            raise ValueError("zero aspect")
        g = math.gcd(width, height)
        self.width, self.height = width // g, height // g

    def as_f32(self):
        return f32(self.width) / f32(self.height)

    def __eq__(self, o):
        return (self.width, self.height) == (o.width, o.height)


AspectRatio.SQUARE = AspectRatio(1, 1)


class Resolution:
    """resolution.rs:8-60."""

    def __init__(self, width: int, height: int):
        self.width, self.height = int(width), int(height)

    def aspect_ratio(self):
        if self.width == 0 or self.height == 0:
            return None
        return AspectRatio(self.width, self.height)

    def fit_aspect_ratio(self, ratio: AspectRatio):
        """resolution.rs:66-106."""
        to = self.aspect_ratio()
        if to is None:
            return Rect.from_top_left(0.0, 0.0, f32(self.width), f32(self.height))
        from_ratio, to_ratio = ratio.as_f32(), to.as_f32()
        if from_ratio > to_ratio:
            w = f32(self.width)
            h = f32(self.width) / from_ratio
            x_min = f32(0.0)
            y_min = (f32(self.height) - h) / f32(2.0)
        else:
            w = f32(self.height) * from_ratio
            h = f32(self.height)
            x_min = (f32(self.width) - w) / f32(2.0)
            y_min = f32(0.0)
        return Rect.from_top_left(x_min, y_min, w, h)

    def __eq__(self, o):
        return (self.width, self.height) == (o.width, o.height)

    def __repr__(self):
        return f"{self.width}x{self.height}"


class Rect:
    """rect.rs:11-237: centre + size, f32."""

    __slots__ = ("cx", "cy", "w", "h")

    def __init__(self, cx, cy, w, h):
        self.cx, self.cy, self.w, self.h = f32(cx), f32(cy), f32(w), f32(h)

    @staticmethod
    def from_center(cx, cy, w, h):
        return Rect(cx, cy, w, h)

    @staticmethod
    def from_top_left(x, y, w, h):
        x, y, w, h = f32(x), f32(y), f32(w), f32(h)
        return Rect(x + w * f32(0.5), y + h * f32(0.5), w, h)

    @staticmethod
    def from_ranges(x0, x1, y0, y1):
        return Rect._span_inner(x0, y0, x1, y1)

    @staticmethod
    def _span_inner(x_min, y_min, x_max, y_max):
        x_min, y_min, x_max, y_max = f32(x_min), f32(y_min), f32(x_max), f32(y_max)
        assert x_min <= x_max and y_min <= y_max
        return Rect.from_top_left(x_min, y_min, x_max - x_min, y_max - y_min)

    @staticmethod
    def bounding(points):
        pts = [(f32(p[0]), f32(p[1])) for p in points]
        if not pts:
            return None
        minx, miny = pts[0]
        maxx, maxy = pts[0]
        for x, y in pts[1:]:
            minx, miny = min(minx, x), min(miny, y)
            maxx, maxy = max(maxx, x), max(maxy, y)
        return Rect._span_inner(minx, miny, maxx, maxy)

    def scale(self, s):
        s = f32(s)
        return Rect(self.cx, self.cy, self.w * s, self.h * s)

    def grow_rel(self, amount):
        """rect.rs:84-94."""
        a = f32(amount)
        left = self.w * a
        right = self.w * a
        top = self.h * a
        bottom = self.h * a
        return Rect(self.cx, self.cy, self.w + left + right, self.h + top + bottom)

    def grow_to_fit_aspect(self, target: AspectRatio):
        """rect.rs:104-117."""
        w, h = self.w, self.h
        target_width = self.h * target.as_f32()
        if target_width >= self.w:
            inc_w = target_width - self.w
            w = w + inc_w
        else:
            target_height = self.w / target.as_f32()
            inc_h = target_height - self.h
            h = h + inc_h
        return Rect(self.cx, self.cy, w, h)

    def grow_move_center(self, x_center, y_center):
        xc, yc = f32(x_center), f32(y_center)
        w = max(abs(xc - self.x()), abs(xc - (self.x() + self.w))) * f32(2.0)
        h = max(abs(yc - self.y()), abs(yc - (self.y() + self.h))) * f32(2.0)
        return Rect(xc, yc, w, h)

    def top_left(self):
        return (self.cx - self.w * f32(0.5), self.cy - self.h * f32(0.5))

    def x(self):
        return self.top_left()[0]

    def y(self):
        return self.top_left()[1]

    def width(self):
        return self.w

    def height(self):
        return self.h

    def area(self):
        return self.w * self.h

    def center(self):
        return (self.cx, self.cy)

    def size(self):
        return (self.w, self.h)

    def move_by(self, off):
        return Rect(self.cx + f32(off[0]), self.cy + f32(off[1]), self.w, self.h)

    def move_to(self, x, y):
        return Rect.from_top_left(x, y, self.w, self.h)

    def intersection(self, other):
        """rect.rs:193-202."""
        ax, ay = self.top_left()
        bx, by = other.top_left()
        minx, miny = max(ax, bx), max(ay, by)
        maxx = min(ax + self.w, bx + other.w)
        maxy = min(ay + self.h, by + other.h)
        if minx > maxx or miny > maxy:
            return None
        return Rect.bounding([(minx, miny), (maxx, maxy)])

    def intersection_area(self, other):
        r = self.intersection(other)
        return f32(0.0) if r is None else r.area()

    def union_area(self, other):
        return self.area() + other.area() - self.intersection_area(other)

    def iou(self, other):
        with np.errstate(divide="ignore", invalid="ignore"):
            return self.intersection_area(other) / self.union_area(other)

    def contains_point(self, p):
        px, py = f32(p[0]), f32(p[1])
        return bool(self.x() <= px and self.y() <= py and self.x() + self.w >= px and self.y() + self.h >= py)

    def corners(self):
        x, y, w, h = self.x(), self.y(), self.w, self.h
        return [(x, y), (x + w, y), (x + w, y + h), (x, y + h)]

    def __eq__(self, o):
        return (self.cx, self.cy, self.w, self.h) == (o.cx, o.cy, o.w, o.h)

    def __repr__(self):
        return f"Rect @ ({self.cx},{self.cy})/{self.w}x{self.h}"

    def as_tuple(self):
        return (self.cx, self.cy, self.w, self.h)


class RotatedRect:
    """rect.rs:269-424."""

    __slots__ = ("rect", "radians")

    def __init__(self, rect: Rect, radians=0.0):
        self.rect, self.radians = rect, f32(radians)

    @staticmethod
    def of(r):
        return r if isinstance(r, RotatedRect) else RotatedRect(r, 0.0)

    @staticmethod
    def bounding(radians, points):
        """rect.rs:287-325."""
        pts = [(f32(p[0]), f32(p[1])) for p in points]
        if not pts:
            return None
        cw = rot_ccw(-f32(radians))
        minx = miny = f32(np.finfo(np.float32).max)
        maxx = maxy = f32(np.finfo(np.float32).min)
        for x, y in pts:
            px, py = mat_mul_vec(cw, x, y)
            minx, miny = min(minx, px), min(miny, py)
            maxx, maxy = max(maxx, px), max(maxy, py)
        cx = (minx + maxx) * f32(0.5)
        cy = (miny + maxy) * f32(0.5)
        cx, cy = rotate_ccw(cx, cy, radians)
        return RotatedRect(Rect.from_center(cx, cy, maxx - minx, maxy - miny), radians)

    def rotation_radians(self):
        return self.radians

    def map(self, f):
        return RotatedRect(f(self.rect), self.radians)

    def center(self):
        return self.rect.center()

    def grow_rel(self, amount):
        return self.map(lambda r: r.grow_rel(amount))

    def grow_to_fit_aspect(self, target):
        return self.map(lambda r: r.grow_to_fit_aspect(target))

    def transform_in(self, pt):
        """rect.rs:405-409."""
        cx, cy = self.rect.w * f32(0.5), self.rect.h * f32(0.5)
        tlx, tly = self.rect.top_left()
        px = f32(pt[0]) - tlx - cx
        py = f32(pt[1]) - tly - cy
        rx, ry = rotate_cw(px, py, self.radians)
        return (rx + cx, ry + cy)

    def transform_out(self, pt):
        """rect.rs:417-423."""
        cx, cy = self.rect.w * f32(0.5), self.rect.h * f32(0.5)
        rx, ry = rotate_ccw(f32(pt[0]) - cx, f32(pt[1]) - cy, self.radians)
        tlx, tly = self.rect.top_left()
        return (rx + cx + tlx, ry + cy + tly)

    def contains_point(self, pt):
        p = self.transform_in(pt)
        return self.rect.move_to(0.0, 0.0).contains_point(p)

    def rotated_corners(self):
        rot = rot_ccw(self.radians)
        out = []
        for (x, y) in self.rect.corners():
            rx, ry = mat_mul_vec(rot, x - self.rect.cx, y - self.rect.cy)
            out.append((self.rect.cx + rx, self.rect.cy + ry))
        return out

    def __eq__(self, o):
        return self.rect == o.rect and self.radians == o.radians

    def __repr__(self):
        return f"RotatedRect({self.rect!r}, {self.radians})"
