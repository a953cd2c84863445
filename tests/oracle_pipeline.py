"""Oracle composition of the per-frame face pipeline (TEST INFRASTRUCTURE): exactly what
`zb_face_pipeline_run` claims to do, written with the oracle's Detector / LandmarkTracker."""
import numpy as np

from oracle.detection import Detector, ShortRangeNetwork
from oracle.image import Image
from oracle.landmark import Estimator, FaceMeshV1, LandmarkTracker


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
