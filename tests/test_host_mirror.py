"""Host-side mirror (zaru_b200.rect / image) against the oracle's restatement: identical f32 results."""
import math

import numpy as np

from oracle import geometry as og
from oracle import image as oi
from zaru_b200.image import Image
from zaru_b200.rect import AspectRatio, Rect, Resolution, RotatedRect


def _t(r):
    return tuple(float(v) for v in (*r.center(), *r.size()))


def test_rect_ops_match_oracle():
    rng = np.random.default_rng(0)
    for _ in range(200):
        cx, cy, w, h = rng.uniform(-500, 2500, 2).tolist() + rng.uniform(1, 900, 2).tolist()
        a, b = Rect.from_center(cx, cy, w, h), og.Rect.from_center(cx, cy, w, h)
        for num, den in [(1, 1), (16, 9), (3, 4)]:
            assert _t(a.grow_to_fit_aspect(AspectRatio(num, den))) == tuple(float(v) for v in b.grow_to_fit_aspect(og.AspectRatio(num, den)).as_tuple())
        amt = float(rng.uniform(0, 2))
        assert _t(a.grow_rel(amt)) == tuple(float(v) for v in b.grow_rel(amt).as_tuple())
        rad = float(rng.uniform(-math.pi, math.pi))
        pa = RotatedRect(a, rad).transform_out((3.5, -7.25))
        pb = og.RotatedRect(b, rad).transform_out((3.5, -7.25))
        assert (float(pa[0]), float(pa[1])) == (float(pb[0]), float(pb[1]))


def test_view_composition_matches_oracle():
    rng = np.random.default_rng(1)
    px = np.zeros((720, 1280, 4), np.uint8)
    img, oimg = Image(px), oi.Image(px)
    for _ in range(100):
        cx, cy = rng.uniform(0, 1280), rng.uniform(0, 720)
        w, h = rng.uniform(20, 900, 2)
        rad1, rad2 = rng.uniform(-1, 1, 2)
        v = img.view(RotatedRect(Rect.from_center(cx, cy, w, h), rad1)).view(RotatedRect(Rect.from_center(w / 3, h / 2, w / 2, h / 2), rad2))
        o = oimg.view(og.RotatedRect(og.Rect.from_center(cx, cy, w, h), rad1)).view(og.RotatedRect(og.Rect.from_center(w / 3, h / 2, w / 2, h / 2), rad2))
        r = v.view_rect()
        assert _t(r.rect()) == tuple(float(x) for x in o.data.rect_.rect.as_tuple())
        assert float(r.rotation_radians()) == float(o.data.rect_.radians)
    assert Resolution(1920, 1080).aspect_ratio() == AspectRatio(16, 9)
    assert img.as_view().view_rect() == RotatedRect(Rect.from_top_left(0, 0, 1280, 720), 0.0)


def test_thread_context_and_pipeline_switches_exist_and_fail_loudly_without_a_gpu():
    """Host API added for the multi-threaded ingest and the detection-gated landmark stage: present in the mirror, and
    - like everything else - an error, not a CPU fallback, when there is no CUDA device."""
    import threading

    import pytest
    import torch

    import zaru_b200
    from zaru_b200 import ZaruError
    from zaru_b200.pipeline import FacePipeline, HandPipeline
    assert callable(zaru_b200.thread_context) and callable(zaru_b200.context_key)
    assert callable(FacePipeline.set_dense) and callable(HandPipeline.set_dense)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    seen = []

    def worker():
        try:
            zaru_b200.thread_context()
        except ZaruError as ex:
            seen.append(str(ex))

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert seen and "no CPU fallback" in seen[0]


def test_rotated_rect_bounding_reference_kats_and_oracle_agreement():
    """Port of rect.rs:661-708 (`test_rotated_rect_bounding`) for the host mirror, plus bit-equality with the oracle's
    restatement on random inputs."""
    tau = np.float32(2 * math.pi)
    assert RotatedRect.bounding(0.0, []) is None
    assert RotatedRect.bounding(0.0, [(0, 0), (1, 1)]) == RotatedRect(Rect.from_top_left(0, 0, 1, 1), 0.0)
    assert RotatedRect.bounding(0.0, [(0, 0), (10, 0)]) == RotatedRect(Rect.from_top_left(0, 0, 10, 0), 0.0)
    for rad in (tau / np.float32(2), tau / np.float32(4)):
        r = RotatedRect.bounding(rad, [(0, 0), (1, 1)])
        want = Rect.from_top_left(0, 0, 1, 1)
        assert np.allclose(_t(r.rect()), _t(want), atol=1e-6) and r.rotation_radians() == rad
    assert np.allclose(_t(RotatedRect.bounding(tau / np.float32(4), [(0, 0), (9, 9)]).rect()), _t(Rect.from_top_left(0, 0, 9, 9)), atol=1e-5)
    rng = np.random.default_rng(5)
    for _ in range(200):
        pts = rng.uniform(-300, 2000, (int(rng.integers(1, 9)), 2)).astype(np.float32)
        rad = np.float32(rng.uniform(-math.pi, math.pi))
        a, b = RotatedRect.bounding(rad, pts), og.RotatedRect.bounding(rad, pts)
        assert _t(a.rect()) == tuple(float(v) for v in b.rect.as_tuple())
        assert float(a.rotation_radians()) == float(b.radians)


def test_face_mesh_result_eye_rects_match_oracle():
    """`LandmarkResultV1::rotation_radians / left_eye / right_eye` (mediapipe.rs:146-192) and the V2 twins (:315-344,
    :407-421): the host mirror's accessors against the oracle's, bit for bit."""
    from oracle.landmark import FaceLandmarks, FaceLandmarksV2
    from zaru_b200.landmark import LandmarkResultV1, LandmarkResultV2
    rng = np.random.default_rng(6)
    for cls, ocls, n in ((LandmarkResultV1, FaceLandmarks, 468), (LandmarkResultV2, FaceLandmarksV2, 478)):
        for _ in range(20):
            pos = rng.uniform(0, 1080, (n, 3)).astype(np.float32)
            m = cls(pos.copy(), np.array([0.9, 0.0], np.float32))
            o = ocls()
            o.positions[:] = pos
            assert float(m.rotation_radians()) == float(o.rotation_radians())
            for got, want in ((m.left_eye(), o.left_eye()), (m.right_eye(), o.right_eye())):
                assert _t(got.rect()) == tuple(float(v) for v in want.rect.as_tuple())
                assert float(got.rotation_radians()) == float(want.radians)


def test_timer_mirror_averages_and_resets_like_the_reference():
    """timer.rs:16-97: EMA with alpha 0.3 over the recorded durations, count, `Display` prints and resets."""
    from zaru_b200.timer import Timer
    t = Timer("infer")
    t.record(0.010)
    t.record(0.020)
    want = np.float32(0.3) * np.float32(0.020) + (np.float32(1.0) - np.float32(0.3)) * np.float32(0.010)
    assert str(t) == f"infer: 2x{float(want) * 1000.0:.1f}ms"
    assert str(t) == "infer: 0x0.0ms"
    t.record(0.004)
    assert str(t) == "infer: 1x4.0ms"
