"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star / SURVEY §8d): preprocessing tensor bit-exact; identical
post-NMS detection sets; boxes / keypoints / landmarks within 1e-3 NORMALISED units (pixels of the
network input / input size); scores and flags within 1e-3 absolute; angles within 1e-3 rad.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from tests.tolerances import assert_logits_close

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def zb():
    import zaru_b200
    zaru_b200.load_library()
    zaru_b200.context()
    return zaru_b200


def _oimg(rgba):
    from oracle.image import Image
    return Image(rgba)


# ------------------------------------------------------------------------------------------------
# a1/a2: image -> tensor, bit exact
# ------------------------------------------------------------------------------------------------
def _views_for(img_w, img_h, rng, k):
    from oracle.geometry import Rect, RotatedRect
    out = [RotatedRect(Rect.from_top_left(0, 0, img_w, img_h), 0.0)]
    for _ in range(k):
        w, h = rng.uniform(8, 1.4 * img_w), rng.uniform(8, 1.4 * img_h)
        cx, cy = rng.uniform(-0.2 * img_w, 1.2 * img_w), rng.uniform(-0.2 * img_h, 1.2 * img_h)
        out.append(RotatedRect(Rect.from_center(cx, cy, w, h), rng.choice([0.0, 0.0, rng.uniform(-3.2, 3.2), math.pi / 2])))
    return out


@pytest.mark.parametrize("size,net", [((1280, 720), (128, 128)), ((535, 535), (192, 192)), ((97, 61), (64, 64)),
                                      ((1920, 1080), (224, 224))])
def test_preprocess_bit_exact(zb, size, net):
    import ctypes as C
    from oracle.image import image_to_tensor
    from zaru_b200 import _ffi
    from zaru_b200.image import ImageBatch
    from zaru_b200.rect import Resolution
    rng = np.random.default_rng(size[0] * 7 + net[0])
    W, H = size
    px = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    oimg = _oimg(px)
    batch = ImageBatch.from_rgba8(Resolution(W, H), px[None])
    views = _views_for(W, H, rng, 12)
    zviews = (_ffi.zb_view * len(views))(*[_ffi.zb_view(0, float(v.rect.cx), float(v.rect.cy), float(v.rect.w),
                                                      float(v.rect.h), float(v.radians)) for v in views])
    for lo, hi, layout in [(-1.0, 1.0, _ffi.ZB_NCHW), (0.0, 1.0, _ffi.ZB_NHWC)]:
        out = np.empty((len(views), 3, net[1], net[0]) if layout == _ffi.ZB_NCHW else (len(views), net[1], net[0], 3), np.float32)
        _ffi.check(_ffi.lib().zb_preprocess(zb.context(), batch._h, zviews, len(views), net[0], net[1], lo, hi, layout,
                                            out.ctypes.data))
        for i, v in enumerate(views):
            want = image_to_tensor(oimg.view(v), net[0], net[1], lo, hi, "NCHW" if layout == _ffi.ZB_NCHW else "NHWC")[0]
            assert np.array_equal(out[i].view(np.uint32), want.view(np.uint32)), f"view {i} {v} differs"


def test_preprocess_letterbox_worked_example(zb):
    """1080p -> 128x128: rows 0..27 and 100..127 are letterbox (-1.0), source px = (15x, 15y-420)."""
    from zaru_b200.image import Image
    from zaru_b200.detection import ShortRangeNetwork
    from zaru_b200.rect import AspectRatio
    rng = np.random.default_rng(3)
    px = rng.integers(0, 256, size=(1080, 1920, 4), dtype=np.uint8)
    img = Image(px)
    cnn = ShortRangeNetwork().cnn()
    rect = img.rect().grow_to_fit_aspect(AspectRatio.SQUARE)
    t = cnn.tensor(img.view(rect))
    assert t.shape == (1, 3, 128, 128)
    assert (t[0, :, :28, :] == -1.0).all() and (t[0, :, 100:, :] == -1.0).all()
    adjust = (np.float32(1.0) - np.float32(-1.0)) / np.float32(255.0)
    for (x, y) in [(0, 28), (127, 99), (64, 64)]:
        want = px[15 * y - 420, 15 * x, :3].astype(np.float32) * adjust + np.float32(-1.0)
        assert np.array_equal(t[0, :, y, x], want)


# ------------------------------------------------------------------------------------------------
# a3: network forward, all five graphs
# ------------------------------------------------------------------------------------------------
NETS = [("face_detection_short_range", -1.0, 128), ("face_landmark", -1.0, 192), ("iris_landmark", -1.0, 64),
        ("palm_detection_lite", 0.0, 192), ("hand_landmark_lite", 0.0, 224)]


@pytest.mark.parametrize("name,lo,size", NETS)
def test_network_forward_matches_oracle(zb, assets_dir, name, lo, size, sad_linus_cropped):
    from oracle import nn as onn
    from oracle.image import image_to_tensor
    from zaru_b200.nn import NeuralNetwork
    path = os.path.join(assets_dir, "onnx", name + ".onnx")
    net = NeuralNetwork.from_path(path)
    onet = onn.NeuralNetwork(path, backend="cv2")
    assert [s for _, s in net.outputs()] == [[1 if d in (None, 0) else d for d in s] for _, s in onet.outputs()]
    rng = np.random.default_rng(5)
    x = np.empty((5, 3, size, size), np.float32)
    x[0] = image_to_tensor(_oimg(sad_linus_cropped).as_view(), size, size, lo, 1.0)[0]
    x[1:3] = rng.uniform(lo, 1.0, size=(2, 3, size, size))
    # smooth random images (closer to natural statistics than white noise)
    coarse = rng.uniform(lo, 1.0, size=(2, 3, 8, 8)).astype(np.float32)
    x[3:5] = np.repeat(np.repeat(coarse, size // 8, axis=2), size // 8, axis=3)
    got = net.estimate(x)
    want = onet.estimate(x)
    want2 = onet.estimate(x[:1], backend="torch")
    for k, (g, r) in enumerate(zip(got, want)):
        assert g.shape == r.shape
        # raw head tensors are in network-input pixel units (or logits): normalise coordinates by input size
        err = float(np.abs(g - r).max())
        noise = float(np.abs(want2[k] - r[:1]).max())   # oracle-vs-oracle floor on the fixture image
        if g.shape[-1] > 2:
            assert err <= max(TOL * size, 4 * noise), (name, k, err, noise)
        else:       # logits / flags: 4e-3 (= 1e-3 on the score) + 2e-5 |v| for the saturated ones, above the oracle-vs-oracle floor
            assert_logits_close(g, r, extra=4 * noise, what=(name, k, err, noise))


def test_network_forward_chunking_is_invisible(zb, assets_dir):
    from zaru_b200.nn import NeuralNetwork
    net = NeuralNetwork.from_path(os.path.join(assets_dir, "onnx", "face_detection_short_range.onnx"))
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, size=(7, 3, 128, 128)).astype(np.float32)
    net.set_chunk(64)
    a = net.estimate(x)
    net.set_chunk(3)   # ragged last chunk
    b = net.estimate(x)
    net.set_chunk(1)
    c = net.estimate(x)
    for u, v, w in zip(a, b, c):
        assert np.array_equal(u, v) and np.array_equal(u, w)


# ------------------------------------------------------------------------------------------------
# a6-a9: decode + NMS + remap
# ------------------------------------------------------------------------------------------------
def _check_dets(got, want, size, what=""):
    assert len(got) == len(want), f"{what}: {len(got)} detections, oracle has {len(want)}"
    for g, w in zip(got, want):
        assert g.anchor == w.anchor, f"{what}: cluster seeds differ ({g.anchor} vs {w.anchor})"
        gv, wv = g.as_vector(), w.as_vector()
        assert abs(gv[0] - wv[0]) <= TOL, (what, "confidence", gv[0], wv[0])
        assert abs(gv[1] - wv[1]) <= TOL, (what, "angle", gv[1], wv[1])
        assert np.abs(gv[2:] - wv[2:]).max() <= TOL * size, (what, "coords", np.abs(gv[2:] - wv[2:]).max())


@pytest.mark.parametrize("kind", ["face", "palm"])
@pytest.mark.parametrize("mode", ["average", "remove"])
def test_extract_nms_matches_oracle_on_random_heads(zb, kind, mode):
    """Decode+NMS in isolation on synthetic head tensors: exercises dense clusters, ties, both modes."""
    from oracle import detection as od
    from zaru_b200 import detection as zd
    from zaru_b200 import _ffi
    onet = od.ShortRangeNetwork() if kind == "face" else od.PalmLiteNetwork()
    znet = zd.ShortRangeNetwork() if kind == "face" else zd.PalmLiteNetwork()
    A = len(onet.anchors())
    P = onet.num_params
    size = 128 if kind == "face" else 192
    rng = np.random.default_rng(11 if kind == "face" else 12)
    n = 6
    boxes = rng.normal(0, 6, size=(n, A, P)).astype(np.float32)
    boxes[..., 2:4] = rng.uniform(10, 60, size=(n, A, 2))
    scores = rng.normal(-4.0, 2.0, size=(n, A, 1)).astype(np.float32)
    scores[1] = -50.0                                   # empty set
    scores[2, :40] = 3.0                                # exact ties (saturating-free) inside one frame
    scores[3] = rng.normal(0.5, 1.0, size=(A, 1))       # hundreds of candidates
    scores[4, ::7] = 20.0                               # sigmoid saturates to 1.0 -> many equal confidences
    det = zd.Detector(znet, capacity=A)
    det.nms_mut().set_mode(zd.SuppressionMode.Average if mode == "average" else zd.SuppressionMode.Remove)
    got = det.extract(boxes, scores)
    for i in range(n):
        cand = []
        onet.extract([boxes[i:i + 1], scores[i:i + 1]], 0.5, cand)
        nms = od.NonMaxSuppression()
        nms.set_mode(mode)
        want = nms.process(cand)
        _check_dets(got[i], want, size, f"{kind}/{mode}/item{i}")
        if mode == "average" and want:
            # in-tree arithmetic: identical op order; only exp()/atan2() may differ from glibc in the last bit
            gv = np.stack([g.as_vector() for g in got[i]])
            wv = np.stack([w.as_vector() for w in want])
            np.testing.assert_array_max_ulp(gv[:, 2:6], wv[:, 2:6], maxulp=4)
    assert len(got[1]) == 0


def test_extract_threshold_iou_and_capacity(zb):
    from oracle import detection as od
    from zaru_b200 import detection as zd
    from zaru_b200 import _ffi
    rng = np.random.default_rng(13)
    A = 896
    boxes = rng.normal(0, 4, size=(1, A, 16)).astype(np.float32)
    boxes[..., 2:4] = 30.0
    scores = rng.normal(0, 2, size=(1, A, 1)).astype(np.float32)
    det = zd.Detector(zd.ShortRangeNetwork(), capacity=A)
    for thresh, iou in [(0.3, 0.3), (0.7, 0.1), (0.5, 0.0), (0.9, 0.9)]:
        det.set_threshold(thresh)
        det.nms_mut().set_iou_thresh(iou)
        got = det.extract(boxes, scores)[0]
        cand = []
        od.ShortRangeNetwork().extract([boxes, scores], thresh, cand)
        nms = od.NonMaxSuppression()
        nms.set_iou_thresh(iou)
        _check_dets(got, nms.process(cand), 128, f"t{thresh}/iou{iou}")
    small = zd.Detector(zd.ShortRangeNetwork(), capacity=2)
    small.set_threshold(0.5)
    small.nms_mut().set_iou_thresh(0.9)
    with pytest.raises(_ffi.ZaruError) as e:
        small.extract(boxes, scores)
    assert e.value.status == _ffi.ZB_ERR_CAPACITY


def test_detects_face_reference_assertion(zb, sad_linus_full):
    """Port of face/detection.rs:164-173 `detects_face`, plus parity with the oracle."""
    from oracle.detection import Detector as ODetector, ShortRangeNetwork as ONet
    from zaru_b200.detection import Detector, ShortRangeNetwork
    from zaru_b200.image import Image
    det = Detector(ShortRangeNetwork())
    dets = det.detect(Image(sad_linus_full))
    assert len(dets) >= 1, "no detection"
    d = dets[0]
    assert d.confidence() >= 0.8, d.confidence()
    assert abs(math.degrees(float(d.angle()))) < 5.0
    want = ODetector(ONet()).detect(_oimg(sad_linus_full))
    _check_dets(dets, want, 128 * (1280 / 128), "sad_linus")   # remapped coords: tolerance scales with the view


def test_detector_batch_on_synthetic_1080p(zb):
    from oracle.detection import Detector as ODetector, ShortRangeNetwork as ONet
    from zaru_b200 import synth
    from zaru_b200.detection import Detector, ShortRangeNetwork
    from zaru_b200.image import ImageBatch
    from zaru_b200.rect import Resolution
    n = 6
    frames = np.stack([synth.s_face_frame(100 + i)[0] for i in range(n)])
    frames[n - 1] = synth.s_face_frame(7, allow_empty=False)[0]
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
    det = Detector(ShortRangeNetwork())
    got = det.detect_batch(batch, want_raw=True)
    raw_b, raw_s = det.last_raw
    odet = ODetector(ONet())
    total = 0
    for i in range(n):
        want = odet.detect(_oimg(frames[i]))
        margin = float(np.abs(odet.last_raw[1]).min())
        assert_logits_close(raw_s[i], odet.last_raw[1][0], what=i)
        if margin < 1e-2:
            continue   # a logit within 1e-2 of the threshold: set identity is not required
        _check_dets(got[i], want, 128 * 15.0, f"frame{i}")
        total += len(want)
    assert total >= 3


def test_detector_on_views_and_rotations(zb, sad_linus_full):
    from oracle.detection import Detector as ODetector, ShortRangeNetwork as ONet
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from zaru_b200.detection import Detector, ShortRangeNetwork
    from zaru_b200.image import Image
    from zaru_b200.rect import Rect, RotatedRect
    img, oimg = Image(sad_linus_full), _oimg(sad_linus_full)
    det, odet = Detector(ShortRangeNetwork()), ODetector(ONet())
    for (cx, cy, w, h, rad) in [(700, 400, 600, 500, 0.0), (700, 420, 700, 700, 0.2), (640, 360, 1280, 720, -0.15)]:
        got = det.detect(img.view(RotatedRect(Rect.from_center(cx, cy, w, h), rad)))
        want = odet.detect(oimg.view(ORR(ORect.from_center(cx, cy, w, h), rad)))
        _check_dets(got, want, 128 * max(w, h) / 128, f"view{(cx, cy, w, h, rad)}")


# ------------------------------------------------------------------------------------------------
# a10: landmark estimators
# ------------------------------------------------------------------------------------------------
def test_estimates_landmarks_reference_assertions(zb, sad_linus_cropped):
    """Port of mediapipe.rs:603-624 (`estimates_landmarks_upright/rotated/rotated2`) + oracle parity."""
    from oracle.geometry import RotatedRect as ORR, f32
    from oracle.landmark import Estimator as OEst, FaceLandmarks, FaceMeshV1 as ONet
    from zaru_b200.image import Image
    from zaru_b200.landmark import Estimator, FaceMeshV1
    from zaru_b200.rect import RotatedRect
    img, oimg = Image(sad_linus_cropped), _oimg(sad_linus_cropped)
    est = Estimator(FaceMeshV1())
    for deg, expected in [(0.0, 0.0), (10.0, -10.0), (-10.0, 10.0)]:
        rad = float(np.radians(f32(deg)))
        view = img.as_view() if deg == 0.0 else img.view(RotatedRect(img.rect(), rad))
        oview = oimg.as_view() if deg == 0.0 else oimg.view(ORR(oimg.rect(), rad))
        r = est.estimate(view)
        assert r.confidence() > 0.9
        fl = FaceLandmarks()
        fl.positions[:] = r.landmarks().positions()
        for angle in (fl.rotation_radians(), fl.left_eye().radians, fl.right_eye().radians):
            assert abs(math.degrees(float(angle)) - expected) < 5.0
        assert fl.left_eye().center()[0] < fl.right_eye().center()[0]
        want = OEst(ONet()).estimate(oview)
        scale = 535.0 / 192.0
        assert abs(float(r.confidence()) - float(want.face_flag)) <= TOL
        assert np.abs(r.landmarks().positions() - want.positions).max() <= TOL * 192 * scale


@pytest.mark.parametrize("which", ["eye", "hand"])
def test_eye_and_hand_estimators_match_oracle(zb, which, sad_linus_cropped, sad_linus_full):
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from oracle.landmark import Estimator as OEst, EyeNetwork as OEye, HandLiteNetwork as OHand
    from zaru_b200.image import Image
    from zaru_b200.landmark import Estimator, EyeNetwork, HandLiteNetwork
    from zaru_b200.rect import Rect, RotatedRect
    if which == "eye":
        px, net, onet, size = sad_linus_cropped, EyeNetwork(), OEye(), 64
        rois = [(200, 235, 90, 60, 0.0), (335, 235, 96, 64, 0.05), (268, 260, 300, 300, -0.3)]
    else:
        px, net, onet, size = sad_linus_full, HandLiteNetwork(), OHand(), 224
        rois = [(640, 360, 600, 600, 0.0), (640, 360, 500, 400, 0.35), (300, 300, 420, 420, -0.35)]
    img, oimg = Image(px), _oimg(px)
    est = Estimator(net)
    for (cx, cy, w, h, rad) in rois:
        r = est.estimate(img.view(RotatedRect(Rect.from_center(cx, cy, w, h), rad)))
        want = OEst(onet).estimate(oimg.view(ORR(ORect.from_center(cx, cy, w, h), rad)))
        scale = max(w, h) / size
        assert np.abs(r.landmarks().positions() - want.positions).max() <= TOL * size * scale
        if which == "hand":
            assert abs(float(r.presence()) - float(want.presence)) <= TOL
            assert abs(float(r.raw_handedness()) - float(want.raw_handedness)) <= TOL


def test_right_eye_flip_rule(zb, sad_linus_cropped):
    """flip_x: tensor mirrored left-right, x un-mirrored in network coordinates (eye.rs:121-125)."""
    from oracle.image import image_to_tensor
    from oracle.landmark import EyeNetwork as OEye
    from oracle.geometry import AspectRatio, Rect as ORect, f32
    from zaru_b200.image import Image
    from zaru_b200.landmark import Estimator, EyeNetwork
    from zaru_b200.rect import Rect
    img, oimg = Image(sad_linus_cropped), _oimg(sad_linus_cropped)
    roi = (335.0, 235.0, 96.0, 96.0)
    est = Estimator(EyeNetwork())
    view = img.view(Rect.from_center(*roi))
    batch, idx = img.device()
    got = est.estimate_views(batch, [view.to_zb_view(idx)], flip_x=[True])[0].landmarks().positions()
    oview = oimg.view(ORect.from_center(*roi))
    t = image_to_tensor(oview, 64, 64, -1.0, 1.0)[:, :, :, ::-1].copy()
    onet = OEye()
    out = onet.cnn().nn.estimate(t)
    e = onet.new_estimate()
    onet.extract(out, e)
    e.positions[:, 0] = -(e.positions[:, 0] - f32(32.0)) + f32(32.0)
    scale = f32(96.0) / f32(64.0)
    e.positions *= scale
    assert np.abs(got - e.positions).max() <= TOL * 64 * float(scale)


# ------------------------------------------------------------------------------------------------
# full face pipeline (config 4 shape, small batch)
# ------------------------------------------------------------------------------------------------
def test_face_pipeline_matches_oracle(zb):
    from tests.oracle_pipeline import face_pipeline
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    seeds = [200, 201, 202, 203, 204]
    frames = np.stack([synth.s_face_frame(s)[0] for s in seeds] + [synth.s_face_frame(1, allow_empty=True)[0] * 0 + 90])
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
    pipe = FacePipeline()
    res = pipe.run(batch)
    checked = 0
    for i in range(len(frames)):
        dets, lm, flag, view_rect, raw = face_pipeline(frames[i])
        if float(np.abs(raw[1]).min()) < 1e-2:
            continue
        _check_dets(res.detections[i], dets, 128 * 15.0, f"frame{i}")
        if not dets:
            assert res.face_flags[i] == -1.0
            continue
        assert np.allclose(res.rois[i, :4], np.asarray(view_rect.rect.as_tuple(), np.float32), atol=TOL * 128 * 15)
        scale = float(view_rect.rect.w) / 192.0
        assert abs(float(res.face_flags[i]) - float(flag)) <= 2e-3
        assert np.abs(res.landmarks[i] - lm).max() <= TOL * 192 * scale + 15.0 * TOL * 128, i
        checked += 1
    assert checked >= 3
    assert len(res.detections[-1]) == 0 and res.face_flags[-1] == -1.0


def test_pipeline_properties_at_batch(zb):
    """Size-independent properties at a larger batch: permutation equivariance and determinism."""
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.pipeline import FacePipeline
    from zaru_b200.rect import Resolution
    uniq = np.stack([synth.s_face_frame(300 + i)[0] for i in range(8)])
    n = 96
    order = np.random.default_rng(0).integers(0, 8, size=n)
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), uniq[order])
    pipe = FacePipeline()
    a = pipe.run(batch)
    b = pipe.run(batch)
    assert np.array_equal(a.landmarks, b.landmarks) and np.array_equal(a.face_flags, b.face_flags)
    first = {}
    for i, u in enumerate(order):
        if u not in first:
            first[u] = i
            continue
        j = first[u]
        assert len(a.detections[i]) == len(a.detections[j])
        for x, y in zip(a.detections[i], a.detections[j]):
            assert np.array_equal(x.as_vector(), y.as_vector())
        assert np.array_equal(a.landmarks[i], a.landmarks[j])


@pytest.mark.parametrize("netname", ["ShortRangeNetwork", "FullRangeNetwork", "PalmLiteNetwork"])
def test_fused_sampling_matches_sample_then_forward(zb, sad_linus_full, netname):
    """The fused stem samples frames on the fly (separable tables for unrotated views, per-texel math for rotated
    ones, Color::NONE outside the image, zero padding outside the tensor) and convolves on the tensor core with the
    ColorMapper folded into FP16 hi / lo weights (`stem_mma_kernel`).  Feeding the network the tensor produced by
    the bit-exact `zb_preprocess` runs the SIMT f32 stem on the same values: the head tensors must agree to f32
    rounding noise (a sampling or padding slip would move them by orders of magnitude more).  With
    ZB_STEM_MMA=0 both paths run the same conv code and the tensors are bit-identical."""
    from zaru_b200 import detection
    from zaru_b200.detection import Detector
    from zaru_b200.image import Image
    from zaru_b200.rect import AspectRatio, Rect, RotatedRect
    img = Image(sad_linus_full)
    # 5x5 -> 24 with the -1..1 map, 3x3 -> 32 on a 192x192 input, 5x5 -> 32 with the 0..1 map (no validity term)
    network = getattr(detection, netname)
    det = Detector(network())
    cnn = network().cnn()
    batch, idx = img.device()
    views = [img.as_view(), img.view(Rect.from_center(700, 400, 500, 333)),
             img.view(RotatedRect(Rect.from_center(640, 360, 900, 700), 0.3)),
             img.view(Rect.from_center(100, 100, 400, 400))]            # partly outside the image
    simt = os.environ.get("ZB_STEM_MMA") == "0"
    for v in views:
        det.detect_views(batch, [v.to_zb_view(idx)], want_raw=True)
        raw_b, raw_s = det.last_raw
        fit = v.view(v.rect().grow_to_fit_aspect(AspectRatio.SQUARE))
        boxes, scores = cnn.nn.estimate(cnn.tensor(fit))
        if simt:
            assert np.array_equal(raw_b, boxes) and np.array_equal(raw_s, scores)
        else:
            # head tensors reach |x| ~ 130; 2e-5 absolute on the logits is 5e-6 on a score
            assert np.abs(raw_b - boxes).max() <= 2e-5 * (1.0 + np.abs(boxes).max()), np.abs(raw_b - boxes).max()
            assert np.abs(raw_s - scores).max() <= 2e-5 * (1.0 + np.abs(scores).max()), np.abs(raw_s - scores).max()


# ------------------------------------------------------------------------------------------------
# ImageView::to_image and Image::clear on the device
# ------------------------------------------------------------------------------------------------
def test_view_to_image_bit_exact_and_clear(zb, sad_linus_full):
    """`ImageView::to_image` (image/mod.rs:314-325) is the sampler without the resampling step: bit-exact against the
    oracle for plain, oversized, fractional-size and rotated views (the reference's own pixel-level view tests,
    image/tests.rs:71-139, pin the oracle)."""
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from zaru_b200.image import Image, ImageBatch
    from zaru_b200.rect import Rect, RotatedRect, Resolution
    img, oimg = Image(sad_linus_full), _oimg(sad_linus_full)
    for (cx, cy, w, h, rad) in [(640, 360, 1280, 720, 0.0), (300, 200, 101.5, 57.25, 0.0), (-20, 700, 200, 150, 0.0),
                                (700, 400, 333, 222, 0.4), (640, 360, 900, 900, -1.2), (100.25, 99.75, 64, 64, 3.0)]:
        got = img.view(RotatedRect(Rect.from_center(cx, cy, w, h), rad)).to_image()
        want = oimg.view(ORR(ORect.from_center(cx, cy, w, h), rad)).to_image()
        assert got.width() == want.width() and got.height() == want.height()
        assert np.array_equal(got._pixels, want.buf), (cx, cy, w, h, rad)
    # a view of a view composes like the reference (ViewData::view)
    got = img.view(RotatedRect(Rect.from_center(700, 400, 500, 400), 0.3)).view(Rect.from_center(250, 200, 120, 80)).to_image()
    want = oimg.view(ORR(ORect.from_center(700, 400, 500, 400), 0.3)).view(ORect.from_center(250, 200, 120, 80)).to_image()
    assert np.array_equal(got._pixels, want.buf)
    # Image::clear
    frames = np.stack([sad_linus_full, sad_linus_full])
    batch = ImageBatch.from_rgba8(Resolution(1280, 720), frames)
    batch.clear((10, 20, 30, 255), first=1, count=1)
    a = batch.frame(0).as_view().to_image()._pixels
    b = batch.frame(1).as_view().to_image()._pixels
    assert np.array_equal(a, sad_linus_full) and (b == np.array([10, 20, 30, 255], np.uint8)).all()


def test_empty_and_degenerate_inputs(zb, sad_linus_full):
    """Edge cases of the boundary: zero views / frames are a no-op, a view entirely outside the image samples
    Color::NONE everywhere (tensor == lo, like the reference's out-of-bounds reads), a 1x1 image works, NaN view
    parameters do not crash (Rust: NaN casts to 0), and wrong frame indices are rejected."""
    from zaru_b200 import _ffi, context
    from zaru_b200.detection import Detector, ShortRangeNetwork
    from zaru_b200.image import Image
    from zaru_b200.landmark import Estimator, FaceMeshV1
    from zaru_b200.rect import Rect, RotatedRect
    lib = _ffi.lib()
    img = Image(sad_linus_full)
    batch, _ = img.device()
    det, est = Detector(ShortRangeNetwork()), Estimator(FaceMeshV1())
    assert det.detect_views(batch, []) == []
    assert est.estimate_views(batch, []) == []
    out = np.full((1, 3, 16, 16), 7.0, np.float32)
    assert lib.zb_preprocess(context(), batch._h, None, 0, 16, 16, -1.0, 1.0, _ffi.ZB_NCHW, out.ctypes.data) == 0
    assert (out == 7.0).all()
    # a view far outside the image: every sample is Color::NONE -> colour-mapped 0 = lo
    far = img.view(RotatedRect(Rect.from_center(-5000.0, -5000.0, 300.0, 300.0), 0.3))
    t = det._cnn.tensor(far)
    assert (t == -1.0).all()
    assert len(det.detect(far)) == 0
    # 1x1 image
    one = Image(np.array([[[10, 20, 30, 255]]], np.uint8))
    t1 = det._cnn.tensor(one.as_view())
    assert np.isfinite(t1).all() and t1.shape == (1, 3, 128, 128)
    assert len(det.detect(one)) == 0
    # NaN view: no crash, a tensor comes back
    nan_view = (_ffi.zb_view * 1)(_ffi.zb_view(0, float("nan"), 100.0, 50.0, 50.0, 0.0))
    out = np.empty((1, 3, 8, 8), np.float32)
    assert lib.zb_preprocess(context(), batch._h, nan_view, 1, 8, 8, 0.0, 1.0, _ffi.ZB_NCHW, out.ctypes.data) == 0
    assert np.isfinite(out).all()
    # frame index out of range
    bad = (_ffi.zb_view * 1)(_ffi.zb_view(3, 10.0, 10.0, 5.0, 5.0, 0.0))
    assert lib.zb_preprocess(context(), batch._h, bad, 1, 8, 8, 0.0, 1.0, _ffi.ZB_NCHW, out.ctypes.data) == _ffi.ZB_ERR_INVALID_ARGUMENT
    # maximum size: one call takes at most 65535 views (one grid dimension); more is refused with a message, not a launch error
    many = (_ffi.zb_view * 65536)()
    big = np.empty((65536, 3, 2, 2), np.float32)
    rc = lib.zb_preprocess(context(), batch._h, many, 65536, 2, 2, 0.0, 1.0, _ffi.ZB_NCHW, big.ctypes.data)
    assert rc != 0 and b"65535" in lib.zb_last_error()
    for j in range(65535):
        many[j] = _ffi.zb_view(0, 10.0, 10.0, 4.0, 4.0, 0.0)
    assert lib.zb_preprocess(context(), batch._h, many, 65535, 2, 2, 0.0, 1.0, _ffi.ZB_NCHW, big.ctypes.data) == 0
    assert np.isfinite(big[:65535]).all() and (big[0] == big[65534]).all()


@pytest.mark.gpu
def test_page_locked_result_buffers(zb, sad_linus_full):
    """zb_host_alloc / zb_host_free (the mirrors' reusable result arrays): page-locked memory is ordinary host memory to every
    entry point - the same call writes the same bytes into a pinned and into a pageable destination - and it is released."""
    from zaru_b200 import _ffi, _pinned, context
    from zaru_b200.image import Image
    lib = _ffi.lib()
    batch, _ = Image(sad_linus_full).device()
    pageable = np.empty((1, 3, 32, 32), np.float32)
    pinned = _pinned.empty((1, 3, 32, 32), np.float32)
    h, w = sad_linus_full.shape[:2]
    whole = (_ffi.zb_view * 1)(_ffi.zb_view(0, w / 2.0, h / 2.0, float(w), float(h), 0.0))
    for dst in (pageable, pinned):
        assert lib.zb_preprocess(context(), batch._h, whole, 1, 32, 32, -1.0, 1.0, _ffi.ZB_NCHW, dst.ctypes.data) == 0
    assert (pageable == pinned).all()
    views = _pinned.ctypes_array(_ffi.zb_view, 3)
    assert len(views) == 3 and views[2].cx == 0.0
    p = C.c_void_p()
    assert lib.zb_host_alloc(C.c_size_t(0), C.byref(p)) == 0 and not p.value          # zero bytes: no allocation, no error
    assert lib.zb_host_alloc(C.c_size_t(4096), C.byref(p)) == 0 and p.value
    lib.zb_host_free(p)
    lib.zb_host_free(None)                                                            # NULL is a no-op
    del pinned, views


# ------------------------------------------------------------------------------------------------
# zaru_image::blend with linear filtering
# ------------------------------------------------------------------------------------------------
def test_blend_matches_oracle_and_reference_kat(zb, sad_linus_cropped):
    """`zaru_image::blend` (zaru-image/src/blend.rs): the reference's own test (`blend_to_partial_target`, :157-178) on the
    device, then scaling blits (up, down, fractional views, views reaching beyond the source, a 180-degree view = flipped
    corners) against the oracle's restatement: equal up to one 8-bit step (f32 weights on both sides; the sRGB encode is
    evaluated with the device's and the host's `pow`)."""
    from oracle.blend import blend as oblend
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from zaru_b200.image import Image, ImageBatch, blend
    from zaru_b200.rect import Rect, RotatedRect, Resolution
    source = np.empty((3, 3, 4), np.uint8)
    source[:] = (0xAA, 0xBB, 0xCC, 0xDD)
    tb = ImageBatch.from_rgba8(Resolution(2, 1), np.zeros((1, 1, 2, 4), np.uint8))
    target = tb.frame(0)
    blend(target.view(Rect.from_top_left(1.0, 0.0, 1.0, 1.0)), Image(source).view(Rect.from_top_left(1.0, 1.0, 1.0, 1.0)))
    assert target.as_view().to_image()._pixels.reshape(-1).tolist() == [0, 0, 0, 0, 0xAA, 0xBB, 0xCC, 0xDD]
    src = sad_linus_cropped
    simg = Image(src)
    cases = [  # (dest size, dest view (cx, cy, w, h, rad), source view)
        ((90, 60), (45, 30, 90, 60, 0.0), (267.5, 267.5, 535, 535, 0.0)),                 # downscale of the whole image
        ((128, 96), (64, 48, 100.5, 66.25, 0.0), (260, 250, 24, 16, 0.0)),                # upscale of a crop, fractional dest view
        ((64, 64), (32, 32, 64, 64, 0.0), (500, 500, 200, 200, 0.0)),                     # source view reaches beyond the image
        ((100, 50), (50, 25, 80, 40, np.pi), (267, 267, 300, 150, 0.0)),                  # 180-degree dest view: flipped corners
        ((64, 64), (40, 20, 100, 90, 0.0), (100.25, 99.75, 64, 64, 0.0)),                 # dest view larger than the dest image
    ]
    worst = 0
    for (dw, dh), dv, sv in cases:
        base = np.full((dh, dw, 4), 77, np.uint8)
        db = ImageBatch.from_rgba8(Resolution(dw, dh), base[None])
        dimg = db.frame(0)
        blend(dimg.view(RotatedRect(Rect.from_center(*dv[:4]), dv[4])), simg.view(RotatedRect(Rect.from_center(*sv[:4]), sv[4])))
        got = dimg.as_view().to_image()._pixels
        want = base.copy()
        oblend(want, ORR(ORect.from_center(*dv[:4]), dv[4]), src, ORR(ORect.from_center(*sv[:4]), sv[4]))
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert (got != 77).any()
        assert diff.max() <= 1, (dv, sv, int(diff.max()))
        assert (diff > 0).mean() < 1e-3
        worst = max(worst, int(diff.max()))
