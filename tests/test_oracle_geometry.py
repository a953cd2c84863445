"""Oracle vs the reference's own known-answer tests for the in-tree arithmetic.

Ports: crates/zaru-image/src/rect.rs:459-717, resolution.rs:190-226,
crates/zaru/src/image/tests.rs:15-139, crates/zaru/src/nn/mod.rs:724-733,
crates/zaru/src/detection/nms.rs:170-218.
"""
import math

import numpy as np

from oracle.detection import Detection, NonMaxSuppression, calculate_anchors
from oracle.geometry import AspectRatio, Rect, Resolution, RotatedRect, f32, round_half_away, sigmoid
from oracle.image import Image, image_to_tensor
from oracle.nn import ColorMapper

TAU = f32(2.0 * math.pi)


def approx(a, b, tol=1e-6):
    return all(abs(float(x) - float(y)) <= tol for x, y in zip(a, b))


def test_contains_point():
    rect = Rect.from_top_left(-5.0, 5.0, 10.0, 5.0)
    assert rect.contains_point([-5.0, 5.0])
    assert rect.contains_point([-5.0 + 9.0, 5.0 + 4.0])
    assert not rect.contains_point([-5.0 + 11.0, 5.0 + 4.0])
    assert not rect.contains_point([-5.0 + 9.0, 5.0 + 5.0 + 1.0])
    empty = Rect.from_center(0.0, 0.0, 0.0, 0.0)
    assert not empty.contains_point([0.0025, 0.0])
    assert not empty.contains_point([0.0, 1.0])


def test_intersection_and_zero():
    a = Rect.from_ranges(0.0, 10.0, 0.0, 10.0)
    p = Rect.from_ranges(5.0, 5.0, 5.0, 5.0)
    assert a.intersection(p) == p and p.intersection(a) == p
    assert p.intersection_area(Rect.from_ranges(6.0, 10.0, 0.0, 10.0)) == 0.0
    zero = Rect.from_center(0.0, 0.0, 0.0, 0.0)
    also = Rect.from_center(1.0, 0.0, 0.0, 0.0)
    assert zero.area() == 0.0 and also.area() == 0.0
    assert zero.intersection_area(also) == 0.0 and zero.union_area(also) == 0.0


def test_iou_quarter():
    smaller = Rect.from_center(9.0, 9.0, 1.0, 1.0)
    bigger = Rect.from_center(9.0, 9.0, 2.0, 2.0)
    assert smaller.area() == 1.0 and bigger.area() == 4.0
    inter = smaller.intersection(bigger)
    assert inter.center() == smaller.center() and inter.size() == smaller.size()
    assert smaller.intersection_area(bigger) == bigger.intersection_area(smaller) == 1.0
    assert smaller.union_area(bigger) == bigger.union_area(smaller) == 4.0
    assert smaller.iou(bigger) == 0.25 and bigger.iou(smaller) == 0.25


def test_bounding():
    assert Rect.bounding([[0.0, 0.0], [1.0, 1.0], [-1.0, -1.0]]) == Rect.from_center(0.0, 0.0, 2.0, 2.0)
    assert Rect.bounding([[1.0, 1.0], [2.0, 2.0]]) == Rect.from_center(1.5, 1.5, 1.0, 1.0)
    assert Rect.bounding([[0.0, 0.0], [10.0, 0.0]]) == Rect.from_center(5.0, 0.0, 10.0, 0.0)
    assert Rect.bounding([]) is None


def test_fit_aspect():
    sq = AspectRatio.SQUARE
    want = Rect.from_center(10.0, 10.0, 100.0, 100.0)
    assert Rect.from_center(10.0, 10.0, 50.0, 100.0).grow_to_fit_aspect(sq) == want
    assert Rect.from_center(10.0, 10.0, 100.0, 50.0).grow_to_fit_aspect(sq) == want
    assert Rect.from_center(10.0, 10.0, 100.0, 98.0).grow_to_fit_aspect(sq) == want


def test_grow_move_center():
    orig = Rect.from_top_left(0.0, 0.0, 0.0, 0.0)
    assert orig.grow_move_center(0.0, 0.0) == orig
    assert orig.grow_move_center(1.0, 0.0) == Rect.from_top_left(0.0, 0.0, 2.0, 0.0)


def test_rotated_rect_transform():
    null = RotatedRect(Rect.from_top_left(0.0, 0.0, 1.0, 1.0), 0.0)
    assert null.transform_in([0.0, 0.0]) == (0.0, 0.0) and null.transform_out([0.0, 0.0]) == (0.0, 0.0)
    assert null.transform_in([1.0, -1.0]) == (1.0, -1.0) and null.transform_out([1.0, -1.0]) == (1.0, -1.0)
    offset = RotatedRect(Rect.from_top_left(10.0, 20.0, 1.0, 1.0), 0.0)
    assert offset.transform_in([0.0, 0.0]) == (-10.0, -20.0)
    assert offset.transform_in([10.0, 20.0]) == (0.0, 0.0)
    right = RotatedRect(Rect.from_top_left(0.0, 0.0, 1.0, 1.0), TAU / f32(4.0))
    assert right.transform_in([0.5, 0.5]) == (0.5, 0.5) and right.transform_out([0.5, 0.5]) == (0.5, 0.5)
    assert approx(right.transform_in([0.0, 0.0]), (0.0, 1.0))
    assert approx(right.transform_out([0.0, 0.0]), (1.0, 0.0))
    assert approx(right.transform_in([1.0, 0.0]), (0.0, 0.0))
    assert approx(right.transform_out([0.0, -1.0]), (2.0, 0.0))
    rect = RotatedRect(Rect.from_top_left(10.0, 20.0, 1.0, 1.0), TAU / f32(2.0))
    assert approx(rect.transform_in([10.0, 20.0]), (1.0, 1.0), 1e-5)
    assert approx(rect.transform_in([11.0, 21.0]), (0.0, 0.0), 1e-5)
    assert approx(rect.transform_out([0.0, 0.0]), (11.0, 21.0), 1e-5)


def test_rotated_rect_contains_point():
    rect = RotatedRect(Rect.from_top_left(0.0, 0.0, 1.0, 1.0), 1.0)
    assert rect.contains_point([0.5, 0.5])
    assert not rect.contains_point([0.0, 1.5]) and not rect.contains_point([1.0, 1.0])
    rect = RotatedRect(Rect.from_top_left(10.0, 20.0, 100.0, 1.0), TAU / f32(2.0))
    assert not rect.contains_point([9.0, 20.5]) and rect.contains_point([10.0, 20.5])
    assert rect.contains_point([100.0, 20.00005]) and not rect.contains_point([55.0, 21.0])
    rect = RotatedRect(Rect.from_center(10.0, 10.0, 51.0, 1.0), TAU / f32(4.0))
    assert rect.contains_point([10.0, 35.0]) and not rect.contains_point([10.0, 36.0])
    assert rect.contains_point([10.0, -15.0]) and not rect.contains_point([10.0, -16.0])
    assert not rect.contains_point([11.0, 0.0]) and not rect.contains_point([9.0, 0.0])


def test_rotated_rect_bounding():
    assert RotatedRect.bounding(0.0, []) is None
    assert RotatedRect.bounding(0.0, [[0.0, 0.0], [1.0, 1.0]]) == RotatedRect(Rect.from_top_left(0.0, 0.0, 1.0, 1.0), 0.0)
    assert RotatedRect.bounding(0.0, [[0.0, 0.0], [10.0, 0.0]]) == RotatedRect(Rect.from_top_left(0.0, 0.0, 10.0, 0.0), 0.0)
    r = RotatedRect.bounding(TAU / f32(4.0), [[0.0, 0.0], [9.0, 9.0]])
    assert approx(r.rect.as_tuple(), Rect.from_top_left(0.0, 0.0, 9.0, 9.0).as_tuple(), 1e-5)
    r = RotatedRect.bounding(TAU / f32(2.0), [[0.0, 0.0], [1.0, 1.0]])
    assert approx(r.rect.as_tuple(), Rect.from_top_left(0.0, 0.0, 1.0, 1.0).as_tuple(), 1e-6)


def test_corners():
    assert Rect.from_center(1.0, 1.0, 4.0, 2.0).corners() == [(-1.0, 0.0), (3.0, 0.0), (3.0, 2.0), (-1.0, 2.0)]


def test_resolution_aspect():
    assert AspectRatio(1920, 1080) == AspectRatio(16, 9)
    assert Resolution(1920, 1080).aspect_ratio().as_f32() == f32(16.0) / f32(9.0)
    r = Resolution(1920, 1080).fit_aspect_ratio(AspectRatio.SQUARE)
    assert r == Rect.from_top_left(420.0, 0.0, 1080.0, 1080.0)
    assert Resolution(0, 5).aspect_ratio() is None


def test_round_half_away():
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, -0.49999997, 3.2, -3.7], np.float32)
    assert round_half_away(x).tolist() == [1.0, 2.0, 3.0, -1.0, -2.0, 0.0, -0.0, 3.0, -4.0]


# --- image views (crates/zaru/src/image/tests.rs) --------------------------------------------
Y, W, R, G, NONE = (255, 255, 0, 255), (255, 255, 255, 255), (255, 0, 0, 255), (0, 255, 0, 255), (0, 0, 0, 0)


def mkimage(rows):
    return Image(np.array(rows, np.uint8))


def test_view_data():
    image = mkimage([[Y, W, W], [W, R, W], [W, W, W]])
    view = image.as_view().data
    assert view.rect() == Rect.from_top_left(0.0, 0.0, 3.0, 3.0)
    center = view.view(Rect.from_top_left(1.0, 1.0, 1.0, 1.0))
    assert center.rect() == Rect.from_top_left(0.0, 0.0, 1.0, 1.0)
    assert center.rect_ == RotatedRect(Rect.from_top_left(1.0, 1.0, 1.0, 1.0), 0.0)
    tl = center.view(Rect.from_top_left(-1.0, -1.0, 2.0, 2.0))
    assert tl.rect_ == RotatedRect(Rect.from_top_left(0.0, 0.0, 2.0, 2.0), 0.0)
    br = center.view(Rect.from_top_left(0.0, 0.0, 2.0, 2.0))
    assert br.rect_ == RotatedRect(Rect.from_top_left(1.0, 1.0, 2.0, 2.0), 0.0)
    br2 = view.view(Rect.from_top_left(1.0, 1.0, 2.0, 2.0)).view(Rect.from_top_left(1.0, 1.0, 2.0, 2.0))
    assert br2.rect_ == RotatedRect(Rect.from_top_left(2.0, 2.0, 2.0, 2.0), 0.0)


def test_rotated_views():
    image = mkimage([[Y, W], [W, R]])
    full = Rect.from_top_left(0.0, 0.0, 2.0, 2.0)
    v = image.view(RotatedRect(full, 0.0))
    assert [v.get(0, 0), v.get(1, 0), v.get(0, 1), v.get(1, 1)] == [Y, W, W, R]
    v = image.view(RotatedRect(full, TAU / f32(2.0)))
    assert [v.get(0, 0), v.get(1, 0), v.get(0, 1), v.get(1, 1)] == [R, W, W, Y]
    right = image.view(RotatedRect(full, TAU / f32(4.0)))
    assert [right.get(0, 0), right.get(1, 0), right.get(0, 1), right.get(1, 1)] == [W, R, Y, W]
    flip = right.view(RotatedRect(full, TAU / f32(4.0)))
    assert [flip.get(0, 0), flip.get(1, 0), flip.get(0, 1), flip.get(1, 1)] == [R, W, W, Y]
    bot_right = right.view(RotatedRect(Rect.from_top_left(-1.0, 1.0, 2.0, 2.0), 0.0))
    assert bot_right.get(0, 0) == NONE and bot_right.get(1, 0) == Y


def test_view_out_of_image():
    image = mkimage([[R, G]])
    view = image.view(Rect.bounding([[1.0, 0.0], [2.0, 1.0]]))
    assert view.rect().w == 1.0 and view.rect().h == 1.0 and view.get(0, 0) == G
    view = image.view(Rect.bounding([[1.0, 0.0], [100.0, 100.0]]))
    assert view.rect().w == 99.0 and view.rect().h == 100.0
    assert view.get(0, 0) == G and view.get(0, 1) == NONE and view.get(1, 0) == NONE


def test_color_mapper():
    m = ColorMapper.linear(-1.0, 1.0)
    assert m.map((0, 0, 0, 255)) == [-1.0, -1.0, -1.0] and m.map((255, 255, 255, 255)) == [1.0, 1.0, 1.0]
    m = ColorMapper.linear(1.0, 2.0)
    assert m.map((0, 0, 0, 255)) == [1.0, 1.0, 1.0] and m.map((255, 255, 255, 255)) == [2.0, 2.0, 2.0]


def test_tensor_layout_and_letterbox():
    # 4x2 image -> 4x4 NCHW tensor through an aspect-fit view: rows 0 and 3 are letterbox (Color::NONE -> lo)
    img = Image(np.arange(4 * 2 * 4, dtype=np.uint8).reshape(2, 4, 4) * 7)
    rect = img.rect().grow_to_fit_aspect(AspectRatio.SQUARE)
    assert rect == Rect.from_center(2.0, 1.0, 4.0, 4.0)
    t = image_to_tensor(img.view(rect), 4, 4, -1.0, 1.0)
    assert t.shape == (1, 3, 4, 4)
    assert (t[0, :, 0, :] == -1.0).all() and (t[0, :, 3, :] == -1.0).all()
    adjust = (f32(1.0) - f32(-1.0)) / f32(255.0)
    assert t[0, 1, 1, 2] == f32(img.buf[0, 2, 1]) * adjust + f32(-1.0)
    nhwc = image_to_tensor(img.view(rect), 4, 4, -1.0, 1.0, layout="NHWC")
    assert (nhwc[0].transpose(2, 0, 1) == t[0]).all()


# --- NMS (crates/zaru/src/detection/nms.rs:170-218) ------------------------------------------
def test_nms_suppresses_non_maximum():
    nms = NonMaxSuppression()
    nms.set_mode("remove")
    rect = Rect.from_center(0.0, 0.0, 1.0, 1.0)
    out = nms.process([Detection(0.6, rect), Detection(0.55, rect.scale(1.5))])
    assert len(out) == 1
    d = out[0]
    assert d.confidence == f32(0.6) and d.rect.center() == (0.0, 0.0) and d.rect.w == 1.0 and d.rect.h == 1.0


def test_nms_ignores_nonoverlapping():
    nms = NonMaxSuppression()
    nms.set_mode("remove")
    out = nms.process([Detection(1.0, Rect.from_center(0.0, 0.0, 1.0, 1.0)), Detection(1.0, Rect.from_center(5.0, 0.0, 1.0, 1.0))])
    assert len(out) == 2


def test_nma_averages_detections():
    nms = NonMaxSuppression()
    nms.set_iou_thresh(0.0)
    rect = Rect.from_center(-1.0, 3.0, 1.0, 1.0)
    out = nms.process([Detection(1.0, rect), Detection(0.5, rect.scale(4.0))])
    assert len(out) == 1
    d = out[0]
    assert d.confidence == 1.0 and d.rect.center() == (-1.0, 3.0) and d.rect.w == 2.0 and d.rect.h == 2.0


def test_anchors():
    a = calculate_anchors([(2, 16, 16), (6, 8, 8)])
    assert len(a) == 896
    assert a[0] == a[1] == (f32(0.5) / f32(16), f32(0.5) / f32(16))
    assert a[2] == (f32(1.5) / f32(16), f32(0.5) / f32(16))
    assert a[512] == (f32(0.5) / f32(8), f32(0.5) / f32(8)) and a[512 + 6][0] == f32(1.5) / f32(8)
    assert len(calculate_anchors([(2, 24, 24), (6, 12, 12)])) == 2016


def test_sigmoid_threshold_edge():
    # conf = 0.5 is KEPT at thresh 0.5 (`conf < thresh` skips), and tiny negative logits round to 0.5
    assert sigmoid(0.0) == f32(0.5)
    assert not (sigmoid(f32(-1e-8)) < f32(0.5))
    assert sigmoid(f32(-1e-3)) < f32(0.5)


def test_blend_reference_kat_and_properties():
    """Port of `blend_to_partial_target` (zaru-image/src/blend.rs:157-178) for the oracle's restatement of blend, plus the
    properties any linear-filter blit has: an identity blit copies, a constant source stays constant under scaling, and
    source UVs outside the image write Color::NONE."""
    from oracle.blend import blend, srgb_decode_lut, srgb_encode
    from oracle.geometry import Rect, RotatedRect
    assert np.array_equal(srgb_encode(srgb_decode_lut()), np.arange(256, dtype=np.uint8))      # decode/encode round trip
    source = np.empty((3, 3, 4), np.uint8)
    source[:] = (0xAA, 0xBB, 0xCC, 0xDD)
    target = np.zeros((1, 2, 4), np.uint8)
    blend(target, RotatedRect(Rect.from_top_left(1.0, 0.0, 1.0, 1.0), 0.0), source, RotatedRect(Rect.from_top_left(1.0, 1.0, 1.0, 1.0), 0.0))
    assert target.reshape(-1).tolist() == [0x00, 0x00, 0x00, 0x00, 0xAA, 0xBB, 0xCC, 0xDD]
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (6, 7, 4), dtype=np.uint8)
    dst = np.zeros((6, 7, 4), np.uint8)
    full = RotatedRect(Rect.from_top_left(0.0, 0.0, 7.0, 6.0), 0.0)
    blend(dst, full, src, full)
    assert np.array_equal(dst, src)                                                           # texel centres hit exactly
    big = np.zeros((12, 14, 4), np.uint8)
    const = np.empty((6, 7, 4), np.uint8)
    const[:] = (10, 200, 33, 128)
    blend(big, RotatedRect(Rect.from_top_left(0.0, 0.0, 14.0, 12.0), 0.0), const, full)
    assert (big == np.array([10, 200, 33, 128], np.uint8)).all()
    out = np.full((4, 4, 4), 9, np.uint8)
    blend(out, RotatedRect(Rect.from_top_left(0.0, 0.0, 4.0, 4.0), 0.0), src, RotatedRect(Rect.from_top_left(5.0, 0.0, 4.0, 4.0), 0.0))
    assert (out[:, 2:] == 0).all() and (out[:, :2] != 9).any()       # right half samples beyond the 7-wide source: NONE
