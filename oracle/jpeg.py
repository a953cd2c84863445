"""The pixel half of a baseline JPEG decode, restated in NumPy (TEST INFRASTRUCTURE).

What `zaru_image::jpeg::decode_jpeg` (crates/zaru-image/src/jpeg.rs:107-222) gets from its `turbojpeg` / `mozjpeg`
backends, i.e. libjpeg-turbo's default decompression pipeline; the arithmetic lives in that third-party library (not in
/root/reference), so this restates its published algorithms:
  * jidctint.c  `jpeg_idct_islow`: 13-bit fixed-point constants, column pass then row pass, descale with rounding;
  * jdsample.c  `h2v1_fancy_upsample` / `h2v2_fancy_upsample`: triangle filter, needs downsampled_width > 2;
  * jdcolor.c   `ycc_rgb_convert`: 16-bit fixed-point tables (1.402, 0.34414, 0.71414, 1.772).
PINNED against libjpeg-turbo ITSELF: tests/test_jpeg.py feeds the coefficients our host front end extracts through these
functions and requires bit-equality with cv2.imdecode (OpenCV bundles libjpeg-turbo 3.1) on the reference's baseline
fixture (3rdparty/img/sad_linus_cropped.jpg) and on 4:2:0 / 4:2:2 / grey re-encodings.
"""
from __future__ import annotations

import numpy as np

CONST_BITS, PASS1_BITS = 13, 2
F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = 2446, 3196, 4433, 6270, 7373, 9633
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = 12299, 15137, 16069, 16819, 20995, 25172


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(v, shift):
    """v: [..., 8] int64 along the last axis."""
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F_0_541
    tmp2 = z1 + z3 * (-F_1_847)
    tmp3 = z1 + z2 * F_0_765
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << CONST_BITS
    tmp1 = (z2 - z3) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175
    tmp0, tmp1, tmp2, tmp3 = tmp0 * F_0_298, tmp1 * F_2_053, tmp2 * F_3_072, tmp3 * F_1_501
    z1, z2, z3, z4 = z1 * (-F_0_899), z2 * (-F_2_562), z3 * (-F_1_961) + z5, z4 * (-F_0_390) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
    return _descale(out, shift)


def idct_islow(coef: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """coef [nblocks, 64] int16 (natural order), qt [64] -> samples [nblocks, 8, 8] uint8."""
    x = coef.astype(np.int64).reshape(-1, 8, 8) * qt.astype(np.int64).reshape(1, 8, 8)
    ws = _idct_1d(np.swapaxes(x, 1, 2), CONST_BITS - PASS1_BITS)          # columns: [block, col, row]
    ws = np.swapaxes(ws, 1, 2)                                           # [block, row, col]
    out = _idct_1d(ws, CONST_BITS + PASS1_BITS + 3)
    return np.clip(out + 128, 0, 255).astype(np.uint8)


def plane_from_blocks(blocks: np.ndarray, bw: int, bh: int) -> np.ndarray:
    return blocks.reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def upsample_h2v1(pl: np.ndarray, cw: int, width: int) -> np.ndarray:
    p = pl[:, :cw].astype(np.int32)
    if cw <= 2:
        return np.repeat(p, 2, axis=1)[:, :width].astype(np.uint8)
    prev = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    nxt = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    even = (p * 3 + prev + 1) >> 2
    odd = (p * 3 + nxt + 2) >> 2
    even[:, 0] = p[:, 0]
    odd[:, -1] = p[:, -1]
    out = np.empty((p.shape[0], 2 * cw), np.int32)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out[:, :width].astype(np.uint8)


def upsample_h2v2(pl: np.ndarray, cw: int, ch: int, width: int, height: int) -> np.ndarray:
    p = pl[:ch, :cw].astype(np.int32)
    if cw <= 2:
        return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)[:height, :width].astype(np.uint8)
    above = np.concatenate([p[:1], p[:-1]], axis=0)
    below = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * ch, 2 * cw), np.int32)
    for v, other in ((0, above), (1, below)):
        col = p * 3 + other
        last = np.concatenate([col[:, :1], col[:, :-1]], axis=1)
        nxt = np.concatenate([col[:, 1:], col[:, -1:]], axis=1)
        even = (col * 3 + last + 8) >> 4
        odd = (col * 3 + nxt + 7) >> 4
        even[:, 0] = (col[:, 0] * 4 + 8) >> 4
        odd[:, -1] = (col[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2], out[v::2, 1::2] = even, odd
    return out[:height, :width].astype(np.uint8)


def ycc_to_rgba(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    yy, b, r = y.astype(np.int32), cb.astype(np.int32) - 128, cr.astype(np.int32) - 128
    out = np.empty(y.shape + (4,), np.uint8)
    out[..., 0] = np.clip(yy + ((91881 * r + 32768) >> 16), 0, 255)
    out[..., 1] = np.clip(yy + ((-22554 * b + 32768 - 46802 * r) >> 16), 0, 255)
    out[..., 2] = np.clip(yy + ((116130 * b + 32768) >> 16), 0, 255)
    out[..., 3] = 255
    return out


def decode_from_coefficients(coef: np.ndarray, blocks_w, blocks_h, qtables: np.ndarray, width: int, height: int, ncomp: int, hs: int, vs: int) -> np.ndarray:
    """coef: [total blocks, 64] int16, component-major -> RGBA8 [height, width, 4]."""
    planes, o = [], 0
    for c in range(ncomp):
        nb = blocks_w[c] * blocks_h[c]
        planes.append(plane_from_blocks(idct_islow(coef[o:o + nb], qtables[c]), blocks_w[c], blocks_h[c]))
        o += nb
    y = planes[0][:height, :width]
    if ncomp == 1:
        return ycc_to_rgba(y, np.full_like(y, 128), np.full_like(y, 128))
    cw, ch = (width + hs - 1) // hs, (height + vs - 1) // vs
    if hs == 1 and vs == 1:
        cb, cr = planes[1][:height, :width], planes[2][:height, :width]
    elif vs == 1:
        cb, cr = upsample_h2v1(planes[1][:height], cw, width), upsample_h2v1(planes[2][:height], cw, width)
    else:
        cb, cr = upsample_h2v2(planes[1], cw, ch, width, height), upsample_h2v2(planes[2], cw, ch, width, height)
    return ycc_to_rgba(y, cb, cr)
