"""The fused per-frame face pipeline: detect -> best detection -> RoI -> face mesh, all on device.

Composition taken from the reference's example loop (crates/zaru/examples/facemesh.rs:36-55:
`detector.detect(&image)`, `tracker.set_roi(best.bounding_rect())`) followed by one
`LandmarkTracker::track` step on the same frame (crates/zaru/src/landmark.rs:456-501).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi, _pinned, context
from .detection import Detection, Detections, Detector, ShortRangeNetwork
from .landmark import EyeLandmarks, EyeNetwork, FaceMeshV1, LandmarkTracker


class FacePipelineResult:
    def __init__(self, detections, landmarks, flags, rois):
        self.detections = detections      # list[Detections], frame coordinates
        self.landmarks = landmarks        # [n,L,3] float32, frame coordinates (L = 468 or 478)
        self.face_flags = flags           # [n] sigmoid(face_flag); -1 where no face was detected
        self.rois = rois                  # [n,5]: view_rect (cx, cy, w, h, radians) used for the mesh


class FacePipeline:
    def __init__(self, capacity: int = 16, detector_network=None, landmark_network=None):
        self._det = (detector_network or ShortRangeNetwork()).cnn()
        self._lm = (landmark_network or FaceMeshV1()).cnn()
        self._cap = capacity
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_face_pipeline_create(context(), self._det.nn._h, self._lm.nn._h, C.byref(h)))
        self._h = h
        self._L = _ffi.lib().zb_face_pipeline_num_landmarks(h)     # 468 (FaceMeshV1) or 478 (FaceMeshV2)
        self._bufs = None

    def set_threshold(self, det_thresh=0.5, iou_thresh=0.3, mode=_ffi.ZB_NMS_AVERAGE):
        _ffi.check(_ffi.lib().zb_face_pipeline_set_threshold(self._h, det_thresh, iou_thresh, mode))

    def set_dense(self, dense: bool):
        """dense=True: run the landmark network over every frame instead of only those with a detection (same results)."""
        _ffi.check(_ffi.lib().zb_face_pipeline_set_dense(self._h, 1 if dense else 0))

    def _buffers(self, n):
        if self._bufs is None or self._bufs[0] != n:
            # page-locked: the results of a step come back at the PCIe rate instead of through the driver's bounce buffers
            self._bufs = (n, _pinned.ctypes_array(_ffi.zb_detection, n * self._cap), _pinned.ctypes_array(C.c_int32, n),
                          _pinned.empty((n, self._L, 3), np.float32), _pinned.empty(n, np.float32), _pinned.ctypes_array(_ffi.zb_view, n))
        return self._bufs

    def run_raw(self, batch, n=None):
        """One pass over `n` frames; results land in reusable host buffers (no Python unpacking)."""
        n = len(batch) if n is None else n
        _, dets, counts, lm, flags, rois = self._buffers(n)
        _ffi.check(_ffi.lib().zb_face_pipeline_run(self._h, batch._h, n, dets, counts, self._cap, lm.ctypes.data,
                                                   flags.ctypes.data, rois))
        return dets, counts, lm, flags, rois

    def run(self, batch, n=None) -> FacePipelineResult:
        n = len(batch) if n is None else n
        dets, counts, lm, flags, rois = self.run_raw(batch, n)
        out = [Detections(Detection(dets[i * self._cap + k]) for k in range(min(counts[i], self._cap)))
               for i in range(n)]
        r = np.array([[v.cx, v.cy, v.w, v.h, v.radians] for v in rois], np.float32).reshape(n, 5)
        return FacePipelineResult(out, lm.copy(), flags.copy(), r)

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_face_pipeline_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FaceStreamTracker:
    """The reference's steady-state loop (crates/zaru/examples/facemesh.rs:36-60) over many camera streams:
    every step `tracker.track(frame)`; only where that returns None the detector runs on that stream's frame and
    the best detection (`max_by_key(TotalF32(confidence))`, last maximum) seeds `tracker.set_roi(bounding_rect)`,
    to be tracked from the NEXT frame on.  Detection therefore costs nothing while tracking holds."""

    def __init__(self, streams: int, detector_network=None, landmark_network=None, capacity: int = 16):
        self.detector = Detector(detector_network or ShortRangeNetwork(), capacity=capacity)
        self.tracker = LandmarkTracker(landmark_network or FaceMeshV1(), streams)
        self._n = streams

    def step(self, batch):
        """Returns (tracking results [n] with None where lost, {stream: Detections} for re-detected streams)."""
        results = self.tracker.track(batch)
        lost = [i for i, r in enumerate(results) if r is None]
        redetected = {}
        if lost:
            res = batch.resolution()
            w, h = float(res.width()), float(res.height())
            views = [_ffi.zb_view(i, w * 0.5, h * 0.5, w, h, 0.0) for i in lost]   # the whole frame of each lost stream
            dets = self.detector.detect_views(batch, views)
            ids, rois = [], []
            for i, ds in zip(lost, dets):
                redetected[i] = ds
                best = None
                for d in ds:   # max_by_key: the LAST maximum in iteration order
                    if best is None or _total_key(d.confidence()) >= _total_key(best.confidence()):
                        best = d
                if best is not None:
                    r = best.bounding_rect()
                    ids.append(i)
                    rois.append((r.center()[0], r.center()[1], r.width(), r.height(), 0.0))
            if ids:
                self.tracker.set_rois(ids, rois)
        return results, redetected


def _total_key(x):
    b = int(np.float32(x).view(np.int32))
    return b ^ 0x7FFFFFFF if b < 0 else b


class HandPipelineResult:
    def __init__(self, detections, landmarks, presence, handedness, rois):
        self.detections = detections      # list[Detections] of palms, frame coordinates
        self.landmarks = landmarks        # [n,21,3] float32, frame coordinates
        self.presence = presence          # [n] hand presence (-1 where no palm was detected)
        self.raw_handedness = handedness  # [n]
        self.rois = rois                  # [n,5]: view_rect (cx, cy, w, h, radians) used for the landmark network


class HandPipeline:
    """Palm detection + hand landmarks on device (BASELINE config 3): per frame the best palm seeds
    `RotatedRect(bounding_rect.grow_rel(1.5), det.angle())` (hand/tracking.rs:136, :159) and one
    `LandmarkTracker::track` step of the hand landmark network runs on that rotated view."""

    def __init__(self, capacity: int = 16):
        from .detection import PalmLiteNetwork
        from .landmark import HandLiteNetwork
        self._det = PalmLiteNetwork().cnn()
        self._lm = HandLiteNetwork().cnn()
        self._cap = capacity
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_hand_pipeline_create(context(), self._det.nn._h, self._lm.nn._h, C.byref(h)))
        self._h = h

    def set_threshold(self, det_thresh=0.5, iou_thresh=0.3, mode=_ffi.ZB_NMS_AVERAGE):
        _ffi.check(_ffi.lib().zb_hand_pipeline_set_threshold(self._h, det_thresh, iou_thresh, mode))

    def set_dense(self, dense: bool):
        """dense=True: run the hand landmark network over every frame instead of only those with a palm (same results)."""
        _ffi.check(_ffi.lib().zb_hand_pipeline_set_dense(self._h, 1 if dense else 0))

    def run_raw(self, batch, n=None):
        """One pass over `n` frames; results land in reusable host buffers (no Python unpacking)."""
        n = len(batch) if n is None else n
        if getattr(self, "_bufs", None) is None or self._bufs[0] != n:
            self._bufs = (n, _pinned.ctypes_array(_ffi.zb_detection, n * self._cap), _pinned.ctypes_array(C.c_int32, n),
                          _pinned.empty((n, 21, 3), np.float32), _pinned.empty((n, 2), np.float32), _pinned.ctypes_array(_ffi.zb_view, n))
        _, dets, counts, lm, sc, rois = self._bufs
        _ffi.check(_ffi.lib().zb_hand_pipeline_run(self._h, batch._h, n, dets, counts, self._cap, lm.ctypes.data, sc.ctypes.data, rois))
        return dets, counts, lm, sc, rois

    def run(self, batch, n=None) -> HandPipelineResult:
        n = len(batch) if n is None else n
        dets, counts, lm, sc, rois = self.run_raw(batch, n)
        lm, sc = lm.copy(), sc.copy()
        out = [Detections(Detection(dets[i * self._cap + k]) for k in range(min(counts[i], self._cap))) for i in range(n)]
        r = np.array([[v.cx, v.cy, v.w, v.h, v.radians] for v in rois], np.float32).reshape(n, 5)
        return HandPipelineResult(out, lm, sc[:, 0].copy(), sc[:, 1].copy(), r)

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_hand_pipeline_destroy(self._h)
                self._h = None
        except Exception:
            pass


class FaceIrisResult:
    def __init__(self, face_landmarks, face_flags, face_view_rects, eye_rois, eye_landmarks):
        self.face_landmarks = face_landmarks    # [n,L,3] frame coordinates
        self.face_flags = face_flags            # [n] sigmoid(face_flag)
        self.face_view_rects = face_view_rects  # [n,5] (cx, cy, w, h, radians): the view the mesh ran on
        self.eye_rois = eye_rois                # [n,2,5]: left / right eye RotatedRect after the margin
        self.eye_landmarks = eye_landmarks      # [n,2,76,3] frame coordinates: 5 iris points, then 71 contour points

    def eyes(self, i):
        """(left, right) `EyeLandmarks` of face i."""
        z = np.zeros(2, np.float32)
        return EyeLandmarks(self.eye_landmarks[i, 0], z), EyeLandmarks(self.eye_landmarks[i, 1], z)


class FaceIrisPipeline:
    """BASELINE config 2 on the device: face mesh on each detector crop -> `left_eye()` / `right_eye()`
    (mediapipe.rs:163-192) -> EyeNetwork on the two eye crops, the right one mirrored (eye.rs:24-28, :121-125).
    The reference ships the pieces but no composition; include/zaru_b200.h states the rule."""

    def __init__(self, mesh_network=None, eye_network=None, eye_margin: float = 0.0):
        self._mesh = (mesh_network or FaceMeshV1()).cnn()
        self._iris = (eye_network or EyeNetwork()).cnn()
        h = C.c_void_p()
        _ffi.check(_ffi.lib().zb_face_iris_pipeline_create(context(), self._mesh.nn._h, self._iris.nn._h, C.byref(h)))
        self._h = h
        self._L = _ffi.lib().zb_face_iris_pipeline_num_landmarks(h)
        self._bufs = None
        if eye_margin:
            self.set_eye_margin(eye_margin)

    def set_eye_margin(self, amount: float):
        _ffi.check(_ffi.lib().zb_face_iris_pipeline_set_eye_margin(self._h, float(amount)))

    def _buffers(self, n):
        if self._bufs is None or self._bufs[0] != n:
            self._bufs = (n, _pinned.empty((n, self._L, 3), np.float32), _pinned.empty(n, np.float32), _pinned.ctypes_array(_ffi.zb_view, n),
                          _pinned.ctypes_array(_ffi.zb_view, 2 * n), _pinned.empty((n, 2, 76, 3), np.float32))
        return self._bufs

    def run_raw(self, batch, rois=None, n=None):
        """rois: ctypes array of zb_view (face crops) or None for every whole frame."""
        n = (len(batch) if rois is None else len(rois)) if n is None else n
        _, lm, flags, vr, er, elm = self._buffers(n)
        _ffi.check(_ffi.lib().zb_face_iris_pipeline_run(self._h, batch._h, rois, n, lm.ctypes.data, flags.ctypes.data, vr, er,
                                                        elm.ctypes.data))
        return lm, flags, vr, er, elm

    def run(self, batch, rois=None) -> FaceIrisResult:
        """rois: list of (frame, cx, cy, w, h[, radians]) face crops, or None for every whole frame."""
        arr = None
        if rois is not None:
            arr = (_ffi.zb_view * len(rois))()
            for j, r in enumerate(rois):
                t = tuple(r)
                arr[j] = _ffi.zb_view(int(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4]), float(t[5]) if len(t) > 5 else 0.0)
        lm, flags, vr, er, elm = self.run_raw(batch, arr)
        n = lm.shape[0]
        v5 = lambda v: [v.cx, v.cy, v.w, v.h, v.radians]
        return FaceIrisResult(lm.copy(), flags.copy(), np.array([v5(v) for v in vr], np.float32).reshape(n, 5),
                              np.array([v5(v) for v in er], np.float32).reshape(n, 2, 5), elm.copy())

    def __del__(self):
        try:
            if self._h:
                _ffi.lib().zb_face_iris_pipeline_destroy(self._h)
                self._h = None
        except Exception:
            pass
