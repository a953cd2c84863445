"""Oracle composition of the per-frame face pipeline (TEST INFRASTRUCTURE): exactly what
`zb_face_pipeline_run` claims to do, written with the oracle's Detector / LandmarkTracker."""
import numpy as np

from oracle.detection import Detector, ShortRangeNetwork
from oracle.image import Image
from oracle.geometry import RotatedRect
from oracle.landmark import Estimator, EyeNetwork, FaceMeshV1, LandmarkTracker


def _total_key(x):
    b = int(np.float32(x).view(np.int32))
    return b ^ 0x7FFFFFFF if b < 0 else b


def face_pipeline(frame_rgba: np.ndarray, backend="cv2"):
    """Returns (detections, landmarks [468,3] or None, face_flag or -1, view_rect or None)."""
    img = Image(frame_rgba)
    det = Detector(ShortRangeNetwork(), backend=backend)
    dets = det.detect(img)
    raw = det.last_raw
    if not dets:
        return dets, None, np.float32(-1.0), None, raw
    best = None
    for d in dets:  # max_by_key returns the LAST maximum (facemesh.rs:49-52)
        if best is None or _total_key(d.confidence) >= _total_key(best.confidence):
            best = d
    tracker = LandmarkTracker(Estimator(FaceMeshV1(), backend=backend))
    tracker.loss_thresh = np.float32(-1.0)   # always report; the flag tells whether tracking would be lost
    tracker.set_roi(best.rect)
    view_rect, est, _ = tracker.track(img)
    return dets, est.positions.copy(), est.face_flag, view_rect, raw


def face_iris_pipeline(frame_rgba: np.ndarray, roi, eye_margin=0.0, backend="cv2", mesh_network=None):
    """BASELINE config 2 (what `zb_face_iris_pipeline_run` claims to do), composed from the reference's pieces:
    tracker.set_roi(roi) + one `LandmarkTracker::track` step (landmark.rs:456-501) -> `left_eye()` / `right_eye()`
    (mediapipe.rs:163-192), each `grow_rel(eye_margin)` -> `Estimator::estimate(&image.view(eye))` with EyeNetwork
    (the right eye mirrored, eye.rs:24-28, :121-125) -> `eye.transform_out` into frame coordinates.
    roi = (cx, cy, w, h, radians).  Returns (face positions [L,3], face_flag, view_rect, [left, right] eye
    RotatedRects, eye positions [2,76,3])."""
    from oracle.geometry import Rect
    img = Image(frame_rgba)
    tracker = LandmarkTracker(Estimator(mesh_network or FaceMeshV1(), backend=backend))
    tracker.loss_thresh = np.float32(-1.0)
    tracker.set_roi(RotatedRect(Rect.from_center(*[np.float32(v) for v in roi[:4]]), np.float32(roi[4])))
    view_rect, est, _ = tracker.track(img)
    face = est.positions.copy()
    eyes, eye_pos = [], np.zeros((2, 76, 3), np.float32)
    for side, rect in enumerate((est.left_eye(), est.right_eye())):
        rect = rect.grow_rel(eye_margin) if eye_margin else rect
        eyes.append(rect)
        e = Estimator(EyeNetwork(), backend=backend).estimate(img.view(rect), flip_x=(side == 1))
        for k, p in enumerate(e.positions):
            ox, oy = rect.transform_out((p[0], p[1]))
            eye_pos[side, k] = (ox, oy, p[2])
    return face, est.face_flag, view_rect, eyes, eye_pos


def hand_pipeline(frame_rgba: np.ndarray, thresh=0.5, backend="cv2", dense=False):
    """BASELINE config 3 (what `zb_hand_pipeline_run` claims to do): palm detector on the whole frame -> best palm ->
    RoI = RotatedRect(bounding_rect.grow_rel(1.5), det.angle()) (hand/tracking.rs:136, :159) -> one
    `LandmarkTracker::track` step of the hand landmark network.  dense: frames without a palm still pay for a hand
    inference (on the centre square), like the device pipeline's dense mode - for timing only.
    Returns (detections, positions [21,3] or None, presence or -1, view_rect or None)."""
    from oracle.detection import PalmLiteNetwork
    from oracle.geometry import Rect
    from oracle.landmark import HandLiteNetwork
    img = Image(frame_rgba)
    det = Detector(PalmLiteNetwork(), backend=backend)
    det.thresh = np.float32(thresh)
    dets = det.detect(img)
    tracker = LandmarkTracker(Estimator(HandLiteNetwork(), backend=backend))
    tracker.loss_thresh = np.float32(-1e9)
    if not dets:
        if dense:
            h, w = frame_rgba.shape[:2]
            tracker.set_roi(RotatedRect(Rect.from_center(w / 2, h / 2, min(w, h), min(w, h)), 0.0))
            tracker.track(img)
        return dets, None, np.float32(-1.0), None
    best = None
    for d in dets:
        if best is None or _total_key(d.confidence) >= _total_key(best.confidence):
            best = d
    tracker.set_roi(RotatedRect(best.rect.grow_rel(1.5), best.angle))
    view_rect, est, _ = tracker.track(img)
    return dets, est.positions.copy(), est.presence, view_rect
