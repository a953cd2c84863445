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
