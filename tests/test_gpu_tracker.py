"""GPU parity for SURVEY 8(f) rank 2: `LandmarkTracker` (landmark.rs:361-502) resident on the device, batched over
streams, against the oracle's restatement of one `track()` step.

Free-running trackers diverge slowly by construction (nearest-neighbour sampling is discontinuous in the RoI), so
the per-step comparison is TEACHER-FORCED: before every step the device RoI is set to the oracle's current RoI; the
step itself (view fitting, rotated sampling, network, confidence gate, landmark mapping, RotatedRect::bounding,
grow_rel) must then agree to the usual tolerance.  A free-running run checks the loop stays locked on the face."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def zb():
    import zaru_b200
    zaru_b200.load_library()
    return zaru_b200


def _oimg(rgba):
    from oracle.image import Image
    return Image(rgba)


def _moving_frames(seed, steps):
    """A short synthetic camera sequence: the same face drifting and slowly rotating over a fixed background."""
    from zaru_b200 import synth
    base, meta = synth.s_face_frame(seed, allow_empty=False)
    frames = [base]
    for t in range(1, steps):
        frames.append(np.roll(np.roll(base, 6 * t, axis=1), 3 * t, axis=0))
    return np.stack(frames)


@pytest.mark.parametrize("which", ["v1", "v2"])
def test_track_step_matches_oracle_teacher_forced(zb, which):
    from oracle.detection import Detector as ODetector, ShortRangeNetwork as ODet
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    from oracle.landmark import Estimator as OEst, FaceMeshV1 as OV1, FaceMeshV2 as OV2, LandmarkTracker as OTracker
    from zaru_b200.image import ImageBatch
    from zaru_b200.landmark import FaceMeshV1, FaceMeshV2, LandmarkTracker
    from zaru_b200.rect import Resolution
    steps = 4
    frames = _moving_frames(500, steps)
    net_size = 192 if which == "v1" else 256
    otr = OTracker(OEst(OV1() if which == "v1" else OV2()))
    dets = ODetector(ODet()).detect(_oimg(frames[0]))
    assert dets, "the synthetic frame must contain a detectable face"
    otr.set_roi(dets[0].rect)
    trk = LandmarkTracker(FaceMeshV1() if which == "v1" else FaceMeshV2(), streams=1)
    one = ImageBatch.from_rgba8(Resolution(1920, 1080), frames[:1])
    tracked_steps = 0
    for t in range(steps):
        roi = otr.roi
        assert roi is not None
        trk.set_roi((roi.rect.cx, roi.rect.cy, roi.rect.w, roi.rect.h, roi.radians))
        one.update(frames[t:t + 1])
        want = otr.track(_oimg(frames[t]))
        got = trk.track(one)[0]
        assert (got is None) == (want is None), t
        if want is None:
            break
        view_rect, est, updated = want
        scale = float(view_rect.rect.w) / net_size
        f16_step = 0.125 if which == "v2" else 0.0
        assert np.allclose(got.view_rect()[:4], view_rect.rect.as_tuple(), atol=1e-3) and abs(got.view_rect()[4] - float(view_rect.radians)) <= 1e-6
        assert abs(float(got.estimate().confidence()) - float(est.face_flag)) <= 2e-3
        lim = (TOL * net_size + f16_step) * scale
        assert np.abs(got.estimate().landmarks().positions() - est.positions).max() <= lim, t
        # updated RoI: angle within 1e-3 rad (+ what one f16 step on the eye corners can do), box within the landmark limit
        up = got.updated_roi()
        eye_dist = float(np.hypot(*(est.positions[263, :2] - est.positions[33, :2])))
        assert abs(up[4] - float(updated.radians)) <= 1e-3 + 2 * lim / eye_dist
        assert np.abs(np.asarray(up[:4]) - np.asarray(updated.rect.as_tuple(), np.float32)).max() <= 4 * lim + 1e-3 * float(updated.rect.w)
        # device RoI state = updated.grow_rel(0.3)
        st = trk.rois()[0]
        assert st is not None and abs(st[2] - up[2] * 1.6) <= 1e-3 * up[2] and abs(st[3] - up[3] * 1.6) <= 1e-3 * up[3]
        tracked_steps += 1
    assert tracked_steps >= 3


def test_tracker_loss_seed_and_batch_semantics(zb):
    """None without an RoI; lost (RoI cleared) when the confidence gate fails; streams are independent; the
    reference's steady-state loop (examples/facemesh.rs) re-seeds from the detector and then tracks."""
    from zaru_b200 import synth
    from zaru_b200.image import ImageBatch
    from zaru_b200.landmark import FaceMeshV1, LandmarkTracker
    from zaru_b200.pipeline import FaceStreamTracker
    from zaru_b200.rect import Resolution
    n = 4
    frames = np.stack([synth.s_face_frame(600 + i, allow_empty=False)[0] for i in range(n - 1)] +
                      [np.full((1080, 1920, 4), 90, np.uint8)])          # last stream: no face at all
    frames[-1, ..., 3] = 255
    batch = ImageBatch.from_rgba8(Resolution(1920, 1080), frames)
    trk = LandmarkTracker(FaceMeshV1(), streams=n)
    assert trk.rois() == [None] * n
    assert trk.track(batch) == [None] * n                                  # `let roi = self.roi?;`
    with pytest.raises(Exception):
        trk.set_roi_padding(-0.1)                                          # assert!(padding >= 0.0)
    # an RoI on a faceless region: confidence below the loss threshold -> None and the RoI is cleared
    trk.set_rois([3], [(960.0, 540.0, 300.0, 300.0, 0.0)])
    res = trk.track(batch)
    assert res[3] is None and trk.rois()[3] is None
    # the steady-state loop: step 1 detects (all lost), step 2 tracks the seeded streams
    loop = FaceStreamTracker(n)
    r1, det1 = loop.step(batch)
    assert r1 == [None] * n and sorted(det1) == list(range(n))
    seeded = [i for i in range(n) if len(det1[i])]
    assert len(seeded) >= 2 and 3 not in seeded
    r2, det2 = loop.step(batch)
    held = [i for i in seeded if r2[i] is not None]
    assert len(held) >= 2
    assert all(i in det2 for i in range(n) if r2[i] is None)              # only lost streams were re-detected
    assert all(i not in det2 for i in held)
    for i in held:
        # free-running: the tracked RoI stays on the detected face
        d = max(det1[i], key=lambda x: float(x.confidence()))
        c = d.bounding_rect().center()
        up = r2[i].updated_roi()
        assert abs(up[0] - float(c[0])) <= 0.5 * up[2] and abs(up[1] - float(c[1])) <= 0.5 * up[3]
        assert float(r2[i].estimate().confidence()) >= 0.5
    r3, _ = loop.step(batch)
    for i in held:
        if r3[i] is not None:   # same frame again: the RoI converges instead of drifting
            assert np.abs(np.asarray(r3[i].updated_roi()[:2]) - np.asarray(r2[i].updated_roi()[:2])).max() <= 0.1 * r2[i].updated_roi()[2]


def test_hand_tracker_step_matches_oracle(zb, sad_linus_full):
    """`LandmarkTracker<hand::landmark::LandmarkResult>` as `HandTracker` builds it (hand/tracking.rs:157-163:
    RoI = RotatedRect(palm.bounding_rect().grow_rel(1.5), palm.angle()), padding 0.4).  The reference ships no hand
    fixture, so the step arithmetic (rotated view, presence gate, wrist / middle-MCP angle, bounding, grow_rel) is
    compared on the face fixture with the gate opened; what the hand network sees there is irrelevant to parity."""
    from oracle.landmark import Estimator as OEst, HandLiteNetwork as OHand, LandmarkTracker as OTracker
    from zaru_b200.image import Image
    from zaru_b200.landmark import HandLiteNetwork, LandmarkTracker
    otr = OTracker(OEst(OHand()))
    otr.loss_thresh = np.float32(-1e9)
    otr.set_roi_padding(0.4)
    trk = LandmarkTracker(HandLiteNetwork(), streams=1)
    trk.set_loss_threshold(-1e9)
    trk.set_roi_padding(0.4)
    img = Image(sad_linus_full)
    batch, _ = img.device()
    from oracle.geometry import Rect as ORect, RotatedRect as ORR
    otr.set_roi(ORR(ORect.from_center(700.0, 400.0, 420.0, 420.0), np.float32(0.3)))
    for t in range(3):
        roi = otr.roi
        trk.set_roi((roi.rect.cx, roi.rect.cy, roi.rect.w, roi.rect.h, roi.radians))
        want = otr.track(_oimg(sad_linus_full))
        got = trk.track(batch)[0]
        assert want is not None and got is not None
        view_rect, est, updated = want
        scale = float(view_rect.rect.w) / 224.0
        lim = TOL * 224 * scale
        assert abs(float(got.estimate().presence()) - float(est.presence)) <= 2e-3
        assert np.abs(got.estimate().landmarks().positions() - est.positions).max() <= lim, t
        up = got.updated_roi()
        span = float(np.hypot(*(est.positions[0, :2] - est.positions[9, :2])))
        assert abs(up[4] - float(updated.radians)) <= 1e-3 + 2 * lim / max(span, 1.0)
        assert np.abs(np.asarray(up[:4]) - np.asarray(updated.rect.as_tuple(), np.float32)).max() <= 4 * lim + 1e-3 * float(updated.rect.w)
        st = trk.rois()[0]
        assert abs(st[2] - up[2] * 1.8) <= 1e-3 * up[2]      # grow_rel(0.4): w + 0.4 w + 0.4 w


def test_widened_entry_points_reject_bad_arguments(zb, sad_linus_full):
    """Error behaviour of the rank 2/4 entry points: a status + message, never an abort (SURVEY 8b 'Errors')."""
    import ctypes as C
    from zaru_b200 import _ffi, context
    from zaru_b200.image import Image, ImageBatch
    from zaru_b200.landmark import EyeNetwork, FaceMeshV1, LandmarkTracker
    from zaru_b200.rect import Resolution
    lib = _ffi.lib()
    trk = LandmarkTracker(FaceMeshV1(), streams=2)
    img = Image(sad_linus_full)
    batch1, _ = img.device()
    with pytest.raises(_ffi.ZaruError) as e:          # one frame for a two-stream tracker
        trk.track(batch1)
    assert e.value.status == _ffi.ZB_ERR_INVALID_ARGUMENT
    with pytest.raises(_ffi.ZaruError):               # stream id out of range
        trk.set_rois([2], [(10.0, 10.0, 5.0, 5.0, 0.0)])
    with pytest.raises(_ffi.ZaruError):               # the eye network has no Confidence / angle: not trackable
        LandmarkTracker(EyeNetwork(), streams=1)
    with pytest.raises(_ffi.ZaruError):               # zero streams
        LandmarkTracker(FaceMeshV1(), streams=0)
    # to_image with a non-positive size, clear on memory the library does not own
    out = np.zeros((4, 4, 4), np.uint8)
    view = (_ffi.zb_view * 1)(_ffi.zb_view(0, 2.0, 2.0, 4.0, 4.0, 0.0))
    assert lib.zb_view_to_image(context(), batch1._h, view, 1, 0, 4, out.ctypes.data) == _ffi.ZB_ERR_INVALID_ARGUMENT
    import torch
    dev = torch.zeros((1, 8, 8, 4), dtype=torch.uint8, device="cuda")
    alias = ImageBatch.alias_device(Resolution(8, 8), dev.data_ptr(), 1, keepalive=dev)
    with pytest.raises(_ffi.ZaruError):
        alias.clear((1, 2, 3, 4))
    assert b"cannot be cleared" in lib.zb_last_error()
