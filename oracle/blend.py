"""`zaru_image::blend(&mut dest_view, &src_view)` (oracle; TEST INFRASTRUCTURE).

Follows crates/zaru-image/src/blend.rs:13-32, :44-93 (a 4-vertex triangle strip over the destination view, UVs from the
source view, default `ColorTargetState` = replace), view.rs:81-119 (`uvs` / `clip_corners`: only the transformed top-left
and bottom-right corners are used, so quad and UV rectangle are AXIS-ALIGNED between those two points, rotated views
included), blend.wgsl:27-38 (UV outside [0,1] -> vec4(0)) and gpu.rs:191-205 (sampler: linear mag / min filter, default
address mode ClampToEdge) on `Rgba8UnormSrgb` textures (image.rs): texels are decoded sRGB -> linear before filtering and the
result is encoded back.

PARITY UNPINNED below the one test the reference holds (`blend_to_partial_target`, blend.rs:157-178): the reference's numbers
come from whatever GPU wgpu finds (sub-texel precision of the bilinear weights and the rasteriser's fill rule at exact
edges are implementation-defined).  This restatement fixes them the way the APIs specify them: pixel centres at +0.5,
top-left fill rule, exact f32 weights, sRGB transfer functions of the sRGB standard evaluated in f64.
"""
from __future__ import annotations

import numpy as np

from .geometry import f32


def srgb_decode_lut() -> np.ndarray:
    """u8 sRGB -> linear f32 (256 entries), evaluated in f64 and rounded once."""
    c = np.arange(256, dtype=np.float64) / 255.0
    lin = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    return lin.astype(np.float32)


def srgb_encode(lin) -> np.ndarray:
    """linear f32 in [0,1] -> u8 sRGB: f64 transfer function, round half up."""
    x = np.clip(np.asarray(lin, np.float64), 0.0, 1.0)
    s = np.where(x <= 0.0031308, x * 12.92, 1.055 * x ** (1.0 / 2.4) - 0.055)
    return np.floor(s * 255.0 + 0.5).astype(np.uint8)


def blend(dest_px: np.ndarray, dest_rect, src_px: np.ndarray, src_rect) -> None:
    """In place on dest_px [H,W,4] u8.  dest_rect / src_rect: oracle.geometry.RotatedRect in their image's coordinates
    (`ViewData::rect`)."""
    lut = srgb_decode_lut()
    dh, dw = dest_px.shape[:2]
    sh, sw = src_px.shape[:2]
    d0 = dest_rect.transform_out((f32(0.0), f32(0.0)))
    d1 = dest_rect.transform_out((dest_rect.rect.w, dest_rect.rect.h))
    s0 = src_rect.transform_out((f32(0.0), f32(0.0)))
    s1 = src_rect.transform_out((src_rect.rect.w, src_rect.rect.h))
    dx0, dy0, dx1, dy1 = f32(d0[0]), f32(d0[1]), f32(d1[0]), f32(d1[1])
    sx0, sy0, sx1, sy1 = f32(s0[0]), f32(s0[1]), f32(s1[0]), f32(s1[1])
    if dx0 == dx1 or dy0 == dy1:
        return
    xlo, xhi = min(dx0, dx1), max(dx0, dx1)
    ylo, yhi = min(dy0, dy1), max(dy0, dy1)
    for y in range(max(0, int(np.floor(ylo)) - 1), min(dh, int(np.ceil(yhi)) + 1)):
        cy = f32(y) + f32(0.5)
        if not (ylo <= cy < yhi):                      # top-left fill rule on pixel centres
            continue
        for x in range(max(0, int(np.floor(xlo)) - 1), min(dw, int(np.ceil(xhi)) + 1)):
            cx = f32(x) + f32(0.5)
            if not (xlo <= cx < xhi):
                continue
            tx = (cx - dx0) / (dx1 - dx0)
            ty = (cy - dy0) / (dy1 - dy0)
            px = sx0 + tx * (sx1 - sx0)                # source position in source-image pixels
            py = sy0 + ty * (sy1 - sy0)
            u, v = px / f32(sw), py / f32(sh)
            if u > f32(1.0) or v > f32(1.0) or u < f32(0.0) or v < f32(0.0):
                dest_px[y, x] = 0                      # Color::NONE (blend.wgsl:31-33)
                continue
            fx, fy = px - f32(0.5), py - f32(0.5)
            x0f, y0f = np.floor(fx), np.floor(fy)
            wx, wy = f32(fx - x0f), f32(fy - y0f)
            x0, y0 = int(x0f), int(y0f)
            xa, xb = min(max(x0, 0), sw - 1), min(max(x0 + 1, 0), sw - 1)       # ClampToEdge
            ya, yb = min(max(y0, 0), sh - 1), min(max(y0 + 1, 0), sh - 1)
            out = np.zeros(4, np.float32)
            for c in range(4):
                def tex(yy, xx):
                    t = src_px[yy, xx, c]
                    return lut[t] if c < 3 else f32(t) / f32(255.0)
                top = tex(ya, xa) + wx * (tex(ya, xb) - tex(ya, xa))
                bot = tex(yb, xa) + wx * (tex(yb, xb) - tex(yb, xa))
                out[c] = top + wy * (bot - top)
            dest_px[y, x, :3] = srgb_encode(out[:3])
            dest_px[y, x, 3] = np.uint8(np.floor(np.float64(min(max(out[3], f32(0.0)), f32(1.0))) * 255.0 + 0.5))
